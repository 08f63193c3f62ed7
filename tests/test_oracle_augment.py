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


def test_spaghetti_rule_is_pinned_to_pil():
    """The reference draws the arcs with PIL's ImageDraw.arc (argus/utils.py:252-275). Our rasterisation rule
    (oracle/augment.py::arc_mask) is checked against the real Pillow on random arcs sampled as the reference samples
    them: intersection-over-union above 0.9 (measured 0.92; the rest is edge pixels), and never far from PIL's ink."""
    from PIL import Image, ImageDraw

    H = W = 128
    rng = np.random.default_rng(0)
    inter = union = 0
    far = total = 0
    for _ in range(300):
        x0, y0 = rng.integers(0, W), rng.integers(0, H)
        x1, y1 = rng.integers(x0, W), rng.integers(y0, H)
        a0, a1 = rng.integers(0, 360), rng.integers(0, 360)
        width = int(rng.uniform(1, 5))
        img = Image.new("L", (W, H), 255)
        ImageDraw.Draw(img).arc((x0, y0, x1, y1), a0, a1, fill=0, width=width)
        pil = np.array(img) == 0
        arc = np.array([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0) / 2 + 0.5, (y1 - y0) / 2 + 0.5,
                        np.cos(np.radians(float(a0))), np.sin(np.radians(float(a0))), np.cos(np.radians(float(a1))),
                        np.sin(np.radians(float(a1))), width, (a1 - a0) % 360], dtype=np.float32)
        ours = A.arc_mask(H, W, arc)
        inter += (pil & ours).sum()
        union += (pil | ours).sum()
        # every pixel we paint lies within 2 pixels of PIL's ink (3x3 dilation twice)
        d = pil.copy()
        for _k in range(2):
            p = np.pad(d, 1)
            d = p[:-2, :-2] | p[:-2, 1:-1] | p[:-2, 2:] | p[1:-1, :-2] | p[1:-1, 1:-1] | p[1:-1, 2:] | p[2:, :-2] | p[2:, 1:-1] | p[2:, 2:]
        far += (ours & ~d).sum()
        total += ours.sum()
    assert inter / union > 0.9, inter / union
    assert far / max(total, 1) < 0.02, far / total


def test_spaghetti_sampling_follows_the_reference():
    arcs = A.spaghetti_params(2000, 10, 256, 256, seed=5, step=3)
    cx, cy, rx, ry, width, sweep = arcs[..., 0], arcs[..., 1], arcs[..., 2], arcs[..., 3], arcs[..., 8], arcs[..., 9]
    x0, x1 = cx - (rx - 0.5), cx + (rx - 0.5)
    y0, y1 = cy - (ry - 0.5), cy + (ry - 0.5)
    assert x0.min() >= 0 and x1.max() <= 255 and y0.min() >= 0 and y1.max() <= 255 and (x1 >= x0).all() and (y1 >= y0).all()
    assert abs(x0.mean() - 127.5) < 3 and abs((x1 - x0).mean() - 127.5 / 2) < 3     # x0 ~ U{0..255}, x1 ~ U{x0..255}
    assert set(np.unique(width)) == {1.0, 2.0, 3.0, 4.0}                            # int(U(1, 5))
    assert 0 <= sweep.min() and sweep.max() <= 359
    assert np.array_equal(arcs, A.spaghetti_params(2000, 10, 256, 256, seed=5, step=3))
    assert not np.array_equal(arcs, A.spaghetti_params(2000, 10, 256, 256, seed=5, step=4))
