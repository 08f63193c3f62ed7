"""CPU: the augmentation oracle's parameter sampler and per-op arithmetic (oracle/augment.py)."""
import numpy as np

from oracle import augment as A


def test_parameter_ranges_and_sharing():
    P = A.sample_params(4000, 2, seed=1, step=0)
    assert 0.47 < (P[:, 0] != 1).mean() < 0.53          # planckian p = 0.5 (data.py:67)
    assert 0.47 < (P[:, 7] > 0).mean() < 0.53           # gaussian blur p = 0.5 (data.py:82)
    assert 0.67 < (P[:, 12] != 1).mean() < 0.73         # motion blur p = 0.7 (data.py:85)
    assert P[:, 2].min() >= -0.2 - 1e-6 and P[:, 2].max() <= 1e-6          # brightness (0.8, 1.0) - 1
    assert P[:, 3].min() >= 0.5 - 1e-6 and P[:, 3].max() <= 1.2 + 1e-6     # contrast
    assert P[:, 4].min() >= 0.25 - 1e-6 and P[:, 4].max() <= 1.2 + 1e-6    # saturation
    assert np.abs(P[:, 5]).max() <= 0.1 * 2 * np.pi + 1e-6                 # hue
    sig = P[P[:, 7] > 0, 7]
    assert sig.min() >= 3 and sig.max() <= 8
    assert np.allclose(P[:, 8:17].sum(1), 1, atol=1e-6)
    assert np.all(P[0::2, 2:7] == P[1::2, 2:7])         # same_on_batch=True: both views share the jiggle
    assert len(np.unique(P[:, 6])) == 24
    assert np.array_equal(A.sample_params(8, 2, seed=1, step=0), P[:16])   # pure function of (seed, step, image)
    assert not np.array_equal(A.sample_params(8, 2, seed=1, step=1), P[:16])


def test_hsv_round_trip_and_identity_params():
    rng = np.random.default_rng(0)
    x = rng.random((3, 32, 32)).astype(np.float32)
    h, s, v = A.rgb_to_hsv(x)
    assert np.abs(A.hsv_to_rgb(h, s, v) - x).max() < 2e-6
    p = np.zeros(A.N_PARAMS, dtype=np.float32)
    p[0] = p[1] = 1; p[3] = p[4] = 1; p[6] = 0; p[12] = 1
    u8 = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
    out = A.augment_image(u8, p)
    assert np.abs(out - u8.transpose(2, 0, 1).astype(np.float32) / 255).max() < 2e-6


def test_blurs_preserve_constants_and_motion_kernel_shape():
    x = np.full((3, 16, 16), 0.4, dtype=np.float32)
    assert np.abs(A.gaussian_blur(x, 5.0) - 0.4).max() < 1e-6           # reflect border keeps constants
    k = A.motion_kernel(np.array([0.0], dtype=np.float32), np.array([0.0], dtype=np.float32))[0].reshape(3, 3)
    assert np.allclose(k, [[0, 0, 0], [1 / 3, 1 / 3, 1 / 3], [0, 0, 0]], atol=1e-6)
    y = A.motion_blur(x, k.reshape(9))
    assert np.abs(y[:, :, 1:-1] - 0.4).max() < 1e-6 and y[0, 0, 0] < 0.3  # zero ('constant') border darkens edges
