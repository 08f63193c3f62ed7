"""CPU: the augmentation oracle's parameter sampler and per-op arithmetic (oracle/augment.py), and the pin of the
spaghetti rasteriser (oracle/pil_arc.py) to the real Pillow the reference draws with (argus/utils.py:252-275)."""
import math

import numpy as np
import pytest

from oracle import augment as A


class _AllOn:
    random_erasing = True
    salt_and_pepper = True


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
    assert np.all(P[:, 24:37] == 0)                      # erasing / salt & pepper are default-off (data.py:35,39)
    assert np.array_equal(A.sample_params(8, 2, seed=1, step=0), P[:16])   # pure function of (seed, step, image)
    assert not np.array_equal(A.sample_params(8, 2, seed=1, step=1), P[:16])


def test_erasing_and_noise_parameters_follow_kornia_ranges():
    H, W = 256, 256
    P = A.sample_params(4000, 2, seed=2, step=5, cfg=_AllOn(), H=H, W=W)
    for base, (s_lo, s_hi), (r_lo, r_hi) in ((24, (0.02, 0.1), (2.0, 3.0)), (29, (0.02, 0.05), (0.8, 1.2))):
        on = P[:, base] != 0
        assert 0.47 < on.mean() < 0.53                                            # p = 0.5 (data.py:54,57)
        x, y, w, h = (P[on, base + k] for k in range(1, 5))
        assert (x >= 0).all() and (y >= 0).all() and (x + w <= W).all() and (y + h <= H).all()
        area, ratio = w * h / (H * W), h / w
        assert area.min() > s_lo * 0.9 and area.max() < s_hi * 1.1                # rounding of w, h to pixels
        assert ratio.min() > r_lo * 0.93 and ratio.max() < r_hi * 1.07
        assert np.all(P[~on, base + 1:base + 5] == 0)
    on = P[:, 34] != 0
    assert 0.67 < on.mean() < 0.73                                                # p = 0.7 (data.py:95)
    assert P[on, 35].min() >= 0.01 and P[on, 35].max() <= 0.06 and P[on, 36].min() >= 0.4 and P[on, 36].max() <= 0.6


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


def test_motion_kernel_matches_torch_warp():
    """kornia builds the motion kernel by rotating a 3x3 image with warp_affine(nearest, zeros, align_corners=True) =
    torch's affine_grid + grid_sample on the OpenCV-convention rotation matrix. The closed-form rule of the oracle (and
    of the CUDA sampler) is checked against exactly that torch pipeline."""
    import torch
    import torch.nn.functional as Fn

    rng = np.random.default_rng(3)
    ang = rng.uniform(-35, 35, 500).astype(np.float32)
    ang = ang[np.abs(np.abs(ang) - 30.0) > 0.05]          # |sin| = 0.5 exactly at 30 deg: nearest-neighbour tie
    d = rng.uniform(-0.5, 0.5, ang.shape[0]).astype(np.float32)
    ours = A.motion_kernel(ang, d).reshape(-1, 3, 3)
    dd = (np.clip(d, -1, 1) + 1) / 2
    k = torch.zeros(ang.shape[0], 1, 3, 3)
    k[:, 0, 1, :] = torch.from_numpy(np.stack([dd, np.full_like(dd, 0.5), 1 - dd], -1))
    a = torch.from_numpy(ang).double() * math.pi / 180
    al, be = torch.cos(a), torch.sin(a)
    cx = cy = 1.0
    M = torch.zeros(ang.shape[0], 3, 3, dtype=torch.float64)
    M[:, 0, 0], M[:, 0, 1], M[:, 0, 2] = al, be, (1 - al) * cx - be * cy          # cv2.getRotationMatrix2D
    M[:, 1, 0], M[:, 1, 1], M[:, 1, 2] = -be, al, be * cx + (1 - al) * cy
    M[:, 2, 2] = 1
    N = torch.tensor([[1.0, 0, -1], [0, 1.0, -1], [0, 0, 1]], dtype=torch.float64)   # pixel -> [-1, 1] for size 3
    theta = torch.linalg.inv(N @ M @ torch.linalg.inv(N))[:, :2, :].float()
    grid = Fn.affine_grid(theta, [ang.shape[0], 1, 3, 3], align_corners=True)
    rot = Fn.grid_sample(k, grid, mode="nearest", padding_mode="zeros", align_corners=True)[:, 0]
    rot = rot / rot.sum(dim=(1, 2), keepdim=True)
    assert np.abs(rot.numpy() - ours).max() < 1e-6


def test_plasma_field_is_a_diamond_square_fractal():
    f = A.plasma_field(256, 256, 0.25, 4242)
    assert f.shape == (256, 256) and f.min() >= 0 and f.max() < 1      # convex combinations of U[0,1) draws
    # the coarsest level is the 3x3 seed grid: the four image corners sit (almost) on seed samples
    g = A.plasma_field(256, 256, 1e-6, 4242)                            # no roughness: pure interpolation of the seed
    seed = A.pixel_uniform(4242, 0, *np.mgrid[0:3, 0:3])
    assert abs(g[0, 0] - seed[0, 0]) < 1e-4 and abs(g[128, 128] - seed[1, 1]) < 1e-4 and abs(g[0, 128] - seed[0, 1]) < 1e-4
    # rougher fields have more fine-scale energy
    hi = np.abs(np.diff(A.plasma_field(128, 128, 0.4, 7), axis=1)).mean()
    lo = np.abs(np.diff(A.plasma_field(128, 128, 0.1, 7), axis=1)).mean()
    assert hi > 1.15 * lo
    assert A.plasma_field(128, 64, 0.2, 1).shape == (128, 64)          # non-square: 5 x 3 seed grid
    p = np.zeros(A.N_PARAMS, dtype=np.float32)
    p[17], p[18], p[19], p[20] = 0.3, -0.5, 0.5, 0.25
    frac = A.plasma_shadow_mask(256, 256, p).mean()
    assert 0.0 <= frac <= 1.0
    p[18] = 0
    assert not A.plasma_shadow_mask(64, 64, p).any()


def test_erase_and_salt_pepper_semantics():
    rng = np.random.default_rng(1)
    u8 = rng.integers(1, 255, (64, 64, 3), dtype=np.uint8)
    p = np.zeros(A.N_PARAMS, dtype=np.float32)
    p[0] = p[1] = 1; p[3] = p[4] = 1; p[6] = -1; p[12] = 1
    p[24:29] = [1, 5, 7, 10, 4]          # black 10 x 4 rectangle at (5, 7)
    p[29:34] = [1, 8, 9, 3, 3]           # white 3 x 3 rectangle at (8, 9), applied second
    out = A.augment_image(u8, p)
    assert (out[:, 7:11, 5:8] == 0).all() and (out[:, 9:12, 8:11] == 1).all() and (out[:, 7:9, 8:15] == 0).all()
    ref = u8.transpose(2, 0, 1).astype(np.float32) / 255
    keep = np.ones((64, 64), bool); keep[7:11, 5:15] = False; keep[9:12, 8:11] = False
    assert np.abs(out[:, keep] - ref[:, keep]).max() < 2e-6
    p[24] = p[29] = 0
    p[34:38] = [1, 0.05, 0.5, 0.7]
    out = A.augment_image(u8, p)
    changed = (np.abs(out - ref) > 1e-6).any(0)
    assert 0.03 < changed.mean() < 0.07                                  # amount = 0.05 of the pixels
    assert set(np.unique(out[:, changed])) <= {0.0, 1.0}
    assert (out[0, changed] == out[1, changed]).all() and (out[0, changed] == out[2, changed]).all()


@pytest.mark.parametrize("size,n", [(64, 1500), (256, 600)])
def test_arc_rasteriser_is_pillow_pixel_for_pixel(size, n):
    """The reference draws the arcs with PIL's ImageDraw.arc (argus/utils.py:252-275). oracle/pil_arc.py restates
    Pillow's rasteriser; this test draws random arcs -- sampled as the reference samples them, plus wider strokes --
    with the REAL Pillow installed here and demands identical pixels."""
    from PIL import Image, ImageDraw

    from oracle.pil_arc import arc_mask

    H = W = size
    rng = np.random.default_rng(size)
    for it in range(n):
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0, W)), int(rng.integers(y0, H))
        a0, a1 = int(rng.integers(0, 360)), int(rng.integers(0, 360))
        width = int(rng.uniform(1, 5)) if it % 4 else int(rng.integers(1, 9))
        img = Image.new("L", (W, H), 255)
        ImageDraw.Draw(img).arc((x0, y0, x1, y1), a0, a1, fill=0, width=width)
        pil = np.array(img) == 0
        ours = arc_mask(H, W, (x0, y0, x1, y1), a0, a1, width)
        assert np.array_equal(ours, pil), ((x0, y0, x1, y1), a0, a1, width, int((ours ^ pil).sum()))


def test_oracle_spaghetti_equals_reference_draw_spaghetti():
    """draw_spaghetti itself (PIL path kept in argus_b200/utils.py, a restatement of argus/utils.py:252-275) against the
    oracle rasteriser fed the same sampled arcs."""
    from PIL import Image

    from argus_b200.utils import draw_spaghetti

    H = W = 96
    np.random.seed(5)
    img = draw_spaghetti(Image.new("RGB", (W, H), (200, 150, 100)), n_arcs=10)
    np.random.seed(5)
    arcs = np.zeros((1, 10, A.ARC_FIELDS), dtype=np.float32)
    for k in range(10):
        x0, y0 = np.random.randint(0, W), np.random.randint(0, H)
        x1, y1 = np.random.randint(x0, W), np.random.randint(y0, H)
        s, e = np.random.randint(0, 360), np.random.randint(0, 360)
        arcs[0, k, :7] = [x0, y0, x1, y1, s, e, int(np.random.uniform(1.0, 5.0))]
    ours = A.draw_spaghetti_u8(np.full((1, H, W, 3), (200, 150, 100), dtype=np.uint8), arcs)[0]
    assert np.array_equal(ours, np.array(img))


def test_spaghetti_sampling_follows_the_reference():
    arcs = A.spaghetti_params(2000, 10, 256, 256, seed=5, step=3)
    x0, y0, x1, y1, a0, a1, width = (arcs[..., k] for k in range(7))
    assert x0.min() >= 0 and x1.max() <= 255 and y0.min() >= 0 and y1.max() <= 255 and (x1 >= x0).all() and (y1 >= y0).all()
    assert abs(x0.mean() - 127.5) < 3 and abs((x1 - x0).mean() - 127.5 / 2) < 3     # x0 ~ U{0..255}, x1 ~ U{x0..255}
    assert set(np.unique(width)) == {1.0, 2.0, 3.0, 4.0}                            # int(U(1, 5))
    assert 0 <= a0.min() and a0.max() <= 359 and 0 <= a1.min() and a1.max() <= 359
    assert np.array_equal(arcs, A.spaghetti_params(2000, 10, 256, 256, seed=5, step=3))
    assert not np.array_equal(arcs, A.spaghetti_params(2000, 10, 256, 256, seed=5, step=4))
