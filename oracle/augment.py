"""ORACLE (test infrastructure only): CPU restatement of the reference's augmentation chain.

Reference call sites: /root/reference/argus/data.py:41-103 (Augmentation builds a kornia AugmentationSequential) and
data.py:213-225 (`/255`, then the chain on a (n_cams, 3, H, W) float image pair). Default-ON stages, in order
(data.py:66-92): RandomPlanckianJitter("blackbody") p=.5 -> ColorJiggle(brightness (0.8,1), contrast (0.5,1.2),
saturation (0.25,1.2), hue (-0.1,0.1), same_on_batch=True, p=1) -> RandomGaussianBlur((5,5),(3,8), p=.5)
-> RandomMotionBlur(3, 35deg, 0.5, p=.7) -> RandomPlasmaShadow(roughness (0.1,0.4), intensity (-0.6,0),
quantity (0,0.5), p=1).

The arithmetic lives in the un-vendored third-party dependency **kornia** (`kornia>=0.7.2`,
/root/reference/pyproject.toml:20; not installed, no network). The per-op semantics below restate kornia 0.7.x's
published algorithms (enhance.adjust_*, color.rgb_to_hsv/hsv_to_rgb, filters.gaussian/motion kernels, the
blackbody illuminant table). Two things cannot be reproduced and are OUR frozen spec instead:
  * random numbers: kornia draws from torch's global RNG; here every parameter is a pure function of
    (seed, step, image index, field) through a splitmix64 hash, identical in numpy and in the CUDA kernel;
  * the plasma fractal: kornia's diamond-square consumes the torch RNG recursively; here it is a 6-octave
    value-noise fractal with the same roughness law, normalised to [0,1] per image like kornia's.
No reference test checks an augmented pixel, so: **parity with kornia itself is unpinned**; what is
pinned is GPU == this oracle on identical parameters (tests/test_augment_gpu.py), parameter ranges, and the
reference's only augmentation-related contract: same seed => same result (tests/test_train.py:69-77).
"""
from __future__ import annotations

import numpy as np

N_PARAMS = 24  # floats per image, layout shared with argus_b200/csrc/augment.cu

# kornia.color._planckian "blackbody" table (25 illuminants, 3000 K .. 15000 K), RGB
_BLACKBODY = np.array([
    [0.6743, 0.4029, 0.0013], [0.6281, 0.4241, 0.1665], [0.5919, 0.4372, 0.2513], [0.5623, 0.4457, 0.3154],
    [0.5376, 0.4515, 0.3672], [0.5163, 0.4555, 0.4103], [0.4979, 0.4584, 0.4468], [0.4816, 0.4604, 0.4782],
    [0.4672, 0.4619, 0.5053], [0.4542, 0.4630, 0.5289], [0.4426, 0.4638, 0.5497], [0.4320, 0.4644, 0.5681],
    [0.4223, 0.4648, 0.5844], [0.4135, 0.4651, 0.5990], [0.4054, 0.4653, 0.6121], [0.3980, 0.4654, 0.6239],
    [0.3911, 0.4655, 0.6346], [0.3847, 0.4656, 0.6444], [0.3787, 0.4656, 0.6532], [0.3732, 0.4656, 0.6613],
    [0.3680, 0.4655, 0.6688], [0.3632, 0.4655, 0.6756], [0.3586, 0.4654, 0.6820], [0.3544, 0.4653, 0.6878],
    [0.3503, 0.4653, 0.6933]], dtype=np.float64)
PLANCK_R = (_BLACKBODY[:, 0] / _BLACKBODY[:, 1]).astype(np.float32)
PLANCK_B = (_BLACKBODY[:, 2] / _BLACKBODY[:, 1]).astype(np.float32)

# all 24 orders of the 4 colour operations (0 brightness, 1 contrast, 2 saturation, 3 hue), lexicographic
ORDERS = []
for a in range(4):
    for b in range(4):
        for c in range(4):
            for d in range(4):
                if len({a, b, c, d}) == 4:
                    ORDERS.append((a, b, c, d))

_M64 = (1 << 64) - 1


def hash_u64(seed: int, step: int, image, field) -> np.ndarray:
    """splitmix64 finaliser of a key built from (seed, step, image, field); vectorised over image/field."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed & _M64) * np.uint64(0x9E3779B97F4A7C15)
               + np.uint64(step & _M64) * np.uint64(0xBF58476D1CE4E5B9)
               + np.asarray(image, dtype=np.uint64) * np.uint64(0x94D049BB133111EB)
               + np.asarray(field, dtype=np.uint64))
        z = key
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, step: int, image, field) -> np.ndarray:
    """float32 uniform in [0,1) with 24 random bits."""
    return ((hash_u64(seed, step, image, field) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24))


def _lerp(u, lo, hi):
    return (np.float32(lo) + u * np.float32(hi - lo)).astype(np.float32)


def motion_kernel(angle_deg: np.ndarray, direction: np.ndarray) -> np.ndarray:
    """kornia get_motion_kernel2d(3, angle, direction, mode='nearest'): middle row [d, .5, 1-d] rotated
    anti-clockwise with nearest sampling about the centre (align_corners), zero padded, normalised. -> (n, 9)"""
    n = angle_deg.shape[0]
    d = (np.clip(direction, -1, 1).astype(np.float32) + np.float32(1)) * np.float32(0.5)
    row = np.stack([d, np.full_like(d, 0.5), np.float32(1) - d], -1)  # k[xs + 1]
    a = angle_deg.astype(np.float32) * np.float32(np.pi / 180.0)
    ca, sa = np.cos(a).astype(np.float32), np.sin(a).astype(np.float32)
    out = np.zeros((n, 3, 3), dtype=np.float32)
    for i in range(3):
        for j in range(3):
            x, y = np.float32(j - 1), np.float32(i - 1)
            xs = np.rint(ca * x - sa * y).astype(np.int32)
            ys = np.rint(sa * x + ca * y).astype(np.int32)
            ok = (ys == 0) & (np.abs(xs) <= 1)
            out[:, i, j] = np.where(ok, row[np.arange(n), np.clip(xs + 1, 0, 2)], np.float32(0))
    s = out.sum(axis=(1, 2), keepdims=True, dtype=np.float32)
    return (out / s).reshape(n, 9).astype(np.float32)


def sample_params(n_pairs: int, n_cams: int, seed: int, step: int, cfg=None) -> np.ndarray:
    """(n_pairs*n_cams, N_PARAMS) float32 parameter table. Image index = pair*n_cams + view; the colour-jiggle
    draws use the PAIR index (same_on_batch=True on the (n_cams,3,H,W) mini-batch, data.py:76,224)."""
    c = _cfg(cfg)
    n = n_pairs * n_cams
    img = np.arange(n, dtype=np.uint64)
    pair = img // np.uint64(n_cams) + np.uint64(1 << 32)  # separate key space for per-pair draws
    P = np.zeros((n, N_PARAMS), dtype=np.float32)
    u = lambda who, field: uniform(seed, step, who, field)  # noqa: E731
    # planckian jitter
    if c["planckian_jitter"]:
        apply = u(img, 0) < np.float32(0.5)
        idx = np.minimum((u(img, 1) * np.float32(25)).astype(np.int32), 24)
        P[:, 0] = np.where(apply, PLANCK_R[idx], np.float32(1))
        P[:, 1] = np.where(apply, PLANCK_B[idx], np.float32(1))
    else:
        P[:, 0] = P[:, 1] = 1
    # colour jiggle (shared by the views of a pair)
    if c["color_jiggle"]:
        P[:, 2] = _lerp(u(pair, 2), c["brightness"][0], c["brightness"][1]) - np.float32(1)
        P[:, 3] = _lerp(u(pair, 3), c["contrast"][0], c["contrast"][1])
        P[:, 4] = _lerp(u(pair, 4), c["saturation"][0], c["saturation"][1])
        P[:, 5] = _lerp(u(pair, 5), c["hue"][0], c["hue"][1]) * np.float32(2 * np.pi)
        P[:, 6] = np.minimum((u(pair, 6) * np.float32(24)).astype(np.int32), 23).astype(np.float32)
    else:
        P[:, 2], P[:, 3], P[:, 4], P[:, 5], P[:, 6] = 0, 1, 1, 0, -1
    # gaussian blur: sigma, 0 = off
    if c["blur"]:
        P[:, 7] = np.where(u(img, 7) < np.float32(0.5), _lerp(u(img, 8), 3.0, 8.0), np.float32(0))
    # motion blur: 3x3 kernel (identity = off)
    ident = np.zeros(9, dtype=np.float32)
    ident[4] = 1
    if c["motion_blur"]:
        k = motion_kernel(_lerp(u(img, 10), -35.0, 35.0), _lerp(u(img, 11), -0.5, 0.5))
        P[:, 8:17] = np.where((u(img, 9) < np.float32(0.7))[:, None], k, ident[None])
    else:
        P[:, 8:17] = ident[None]
    # plasma shadow
    if c["plasma_shadow"]:
        P[:, 17] = _lerp(u(img, 12), 0.1, 0.4)
        P[:, 18] = _lerp(u(img, 13), -0.6, 0.0)
        P[:, 19] = _lerp(u(img, 14), 0.0, 0.5)
    else:
        P[:, 17], P[:, 18], P[:, 19] = 0.25, 0.0, 0.0
    P[:, 20] = u(img, 15)  # fractal seed (its 24 random bits, recovered as int(p*2^24))
    return P


def _cfg(cfg):
    d = dict(brightness=(0.8, 1.0), contrast=(0.5, 1.2), saturation=(0.25, 1.2), hue=(-0.1, 0.1),
             color_jiggle=True, planckian_jitter=True, blur=True, motion_blur=True, plasma_shadow=True)
    if cfg is not None:
        for k in d:
            if hasattr(cfg, k):
                v = getattr(cfg, k)
                d[k] = tuple(v) if isinstance(v, (tuple, list)) else v
    return d


# ----------------------------------------------------------------------------------------------------------------
# per-op arithmetic (float32 throughout, like the kernel)
# ----------------------------------------------------------------------------------------------------------------
F = np.float32
TWO_PI = F(2 * np.pi)


def rgb_to_hsv(rgb):
    """kornia.color.rgb_to_hsv: h in [0, 2pi), s, v in [0,1]. rgb (3,H,W) float32."""
    r, g, b = rgb
    mx = np.maximum(np.maximum(r, g), b)
    mn = np.minimum(np.minimum(r, g), b)
    v = mx
    delta = mx - mn
    s = delta / (v + F(1e-8))
    dz = np.where(delta == 0, F(1), delta)
    rc, gc, bc = mx - r, mx - g, mx - b
    # first maximum wins (torch.max returns the first index): r, then g, then b
    h = np.where(r == mx, bc - gc, np.where(g == mx, (rc - bc) + F(2) * dz, (gc - rc) + F(4) * dz))
    h = h / dz
    h = (h / F(6))
    h = h - np.floor(h)
    return TWO_PI * h, s, v


def hsv_to_rgb(h, s, v):
    h6 = (h / TWO_PI) * F(6)
    fl = np.floor(h6)
    hi = np.mod(fl, 6).astype(np.int32)
    f = h6 - fl
    p = v * (F(1) - s)
    q = v * (F(1) - f * s)
    t = v * (F(1) - (F(1) - f) * s)
    r = np.choose(hi, [v, q, p, p, t, v])
    g = np.choose(hi, [t, v, v, q, p, p])
    b = np.choose(hi, [p, p, t, v, v, q])
    return np.stack([r, g, b]).astype(np.float32)


def color_ops(rgb, p):
    """planckian gain then the four jiggle ops in the sampled order. rgb (3,H,W) float32 in [0,1]."""
    x = rgb.astype(np.float32).copy()
    x[0] = np.minimum(x[0] * p[0], F(1))
    x[2] = np.minimum(x[2] * p[1], F(1))
    order = int(p[6])
    if order < 0:
        return x
    for op in ORDERS[order]:
        if op == 0:
            x = np.clip(x + p[2], F(0), F(1))
        elif op == 1:
            x = np.clip(x * p[3], F(0), F(1))
        elif op == 2:
            h, s, v = rgb_to_hsv(x)
            x = hsv_to_rgb(h, np.clip(s * p[4], F(0), F(1)), v)
        else:
            h, s, v = rgb_to_hsv(x)
            h = h + p[5]
            h = h - TWO_PI * np.floor(h / TWO_PI)
            x = hsv_to_rgb(h, s, v)
    return x.astype(np.float32)


def gaussian_taps(sigma):
    k = np.exp(-(np.arange(5, dtype=np.float32) - F(2)) ** 2 / (F(2) * F(sigma) * F(sigma))).astype(np.float32)
    return (k / k.sum(dtype=np.float32)).astype(np.float32)


def gaussian_blur(x, sigma):
    """5x5 separable gaussian, reflect border (kornia default border_type='reflect')."""
    if sigma <= 0:
        return x
    k = gaussian_taps(sigma)
    xp = np.pad(x, ((0, 0), (0, 0), (2, 2)), mode="reflect")
    W = x.shape[2]
    y = sum(k[i] * xp[:, :, i:i + W] for i in range(5)).astype(np.float32)
    yp = np.pad(y, ((0, 0), (2, 2), (0, 0)), mode="reflect")
    H = x.shape[1]
    return sum(k[i] * yp[:, i:i + H, :] for i in range(5)).astype(np.float32)


def motion_blur(x, k9):
    """3x3 correlation with zero ('constant') border."""
    k = k9.reshape(3, 3)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1)))
    H, W = x.shape[1:]
    out = np.zeros_like(x)
    for i in range(3):
        for j in range(3):
            out += k[i, j] * xp[:, i:i + H, j:j + W]
    return out.astype(np.float32)


def _lattice(seed_bits: int, octave: int, iy, ix):
    """hash lattice value in [0,1) for the plasma fractal."""
    key = (np.uint64(seed_bits) << np.uint64(40)) | (np.uint64(octave) << np.uint64(32)) | \
          (np.asarray(iy, dtype=np.uint64) << np.uint64(16)) | np.asarray(ix, dtype=np.uint64)
    return ((hash_u64(0x504C41534D41, 0, key, 0) >> np.uint64(40)).astype(np.float32) * F(2.0 ** -24))


def plasma_field(H, W, roughness, seed_bits):
    """6-octave value noise: octave l has 2^(l+1) cells per side, amplitude roughness^l. (H,W) float32, un-normalised."""
    ys = (np.arange(H, dtype=np.float32) + F(0.5)) / F(H)
    xs = (np.arange(W, dtype=np.float32) + F(0.5)) / F(W)
    field = np.zeros((H, W), dtype=np.float32)
    amp = F(1)
    for l in range(6):
        cells = F(2 << l)
        fy, fx = ys * cells, xs * cells
        iy, ix = np.floor(fy).astype(np.int64), np.floor(fx).astype(np.int64)
        ty, tx = (fy - iy).astype(np.float32), (fx - ix).astype(np.float32)
        ty = ty * ty * (F(3) - F(2) * ty)
        tx = tx * tx * (F(3) - F(2) * tx)
        v00 = _lattice(seed_bits, l, iy[:, None], ix[None, :])
        v01 = _lattice(seed_bits, l, iy[:, None], ix[None, :] + 1)
        v10 = _lattice(seed_bits, l, iy[:, None] + 1, ix[None, :])
        v11 = _lattice(seed_bits, l, iy[:, None] + 1, ix[None, :] + 1)
        top = v00 + (v01 - v00) * tx[None, :]
        bot = v10 + (v11 - v10) * tx[None, :]
        field = field + amp * (top + (bot - top) * ty[:, None])
        amp = amp * F(roughness)
    return field.astype(np.float32)


def augment_image(u8_hwc: np.ndarray, p: np.ndarray) -> np.ndarray:
    """One image: uint8 (H,W,3) -> float32 (3,H,W) in [0,1], parameters p (N_PARAMS,)."""
    x = (u8_hwc.astype(np.float32) * F(1.0 / 255.0)).transpose(2, 0, 1)
    x = color_ops(x, p)
    x = gaussian_blur(x, float(p[7]))
    x = motion_blur(x, p[8:17])
    H, W = x.shape[1:]
    if p[18] != 0:
        f = plasma_field(H, W, p[17], int(round(float(p[20]) * (1 << 24))))
        lo, hi = f.min(), f.max()
        fn = (f - lo) / np.maximum(hi - lo, F(1e-12))
        x = x + np.where(fn < p[19], p[18], F(0))[None]
    return np.clip(x, F(0), F(1)).astype(np.float32)


def augment_batch_u8(images_u8: np.ndarray, seed: int = 0, step: int = 0, cfg=None, params=None) -> np.ndarray:
    """(B, n_cams, H, W, 3) uint8 -> (B, n_cams, 3, H, W) float32 (what Dataset.__getitem__ hands to the model after
    `reshape(-1, H, W)`, data.py:224-227)."""
    B, n_cams, H, W, _ = images_u8.shape
    if params is None:
        params = sample_params(B, n_cams, seed, step, cfg)
    out = np.empty((B, n_cams, 3, H, W), dtype=np.float32)
    for b in range(B):
        for v in range(n_cams):
            out[b, v] = augment_image(images_u8[b, v], params[b * n_cams + v])
    return out


# ------------------------------------------------------------------------------------------------------------------
# spaghetti arcs (reference: argus/utils.py:252-275 `draw_spaghetti`, applied at argus/data.py:212-215 to the decoded
# image BEFORE the kornia chain). The reference draws with PIL's ImageDraw.arc; this restatement is OUR rasterisation
# rule, calibrated against Pillow 12 (tests/test_oracle_augment.py: IoU 0.92 over random arcs, the rest is edge pixels):
#   bbox (x0, y0, x1, y1) -> centre ((x0+x1)/2, (y0+y1)/2), radii ((x1-x0)/2 + 0.5, (y1-y0)/2 + 0.5);
#   a pixel is painted black iff it is inside the outer ellipse, not strictly inside the ellipse shrunk by `width`,
#   and its PARAMETRIC angle atan2(dy/ry, dx/rx) (clockwise from 3 o'clock, as PIL measures) lies in [start, end];
#   the angle test is done with cross products against (cos, sin) of start / end, no transcendental per pixel.
# Sampling follows the reference: x0 ~ U{0..W-1}, y0 ~ U{0..H-1}, x1 ~ U{x0..W-1}, y1 ~ U{y0..H-1},
# start, end ~ U{0..359}, width = int(U(1, 5)); one draw per (seed, step, image, arc, field) through the same hash.
# ------------------------------------------------------------------------------------------------------------------
ARC_FIELDS = 10   # cx, cy, rx, ry, cos0, sin0, cos1, sin1, width, sweep_deg
_ARC_FIELD_BASE = 1000


def spaghetti_params(n_images: int, n_arcs: int, H: int, W: int, seed: int, step: int) -> np.ndarray:
    arcs = np.zeros((n_images, n_arcs, ARC_FIELDS), dtype=np.float32)
    img = np.arange(n_images, dtype=np.uint64)[:, None]
    a = np.arange(n_arcs, dtype=np.uint64)[None, :]

    def U(k):
        return uniform(seed, step, img, np.uint64(_ARC_FIELD_BASE) + a * np.uint64(8) + np.uint64(k))

    def randint(u, lo, hi):   # integer in [lo, hi) from a float32 uniform, lo/hi int arrays
        span = (hi - lo).astype(np.float32)
        v = lo + np.floor(u * span).astype(np.int64)
        return np.minimum(v, hi - 1)

    zero = np.zeros((n_images, n_arcs), dtype=np.int64)
    x0 = randint(U(0), zero, zero + W)
    y0 = randint(U(1), zero, zero + H)
    x1 = randint(U(2), x0, zero + W)
    y1 = randint(U(3), y0, zero + H)
    a0 = randint(U(4), zero, zero + 360)
    a1 = randint(U(5), zero, zero + 360)
    width = np.floor(np.float32(1.0) + U(6) * np.float32(4.0)).astype(np.float32)
    arcs[..., 0] = (x0 + x1).astype(np.float32) * np.float32(0.5)
    arcs[..., 1] = (y0 + y1).astype(np.float32) * np.float32(0.5)
    arcs[..., 2] = (x1 - x0).astype(np.float32) * np.float32(0.5) + np.float32(0.5)
    arcs[..., 3] = (y1 - y0).astype(np.float32) * np.float32(0.5) + np.float32(0.5)
    arcs[..., 4] = np.cos(np.radians(a0.astype(np.float64))).astype(np.float32)
    arcs[..., 5] = np.sin(np.radians(a0.astype(np.float64))).astype(np.float32)
    arcs[..., 6] = np.cos(np.radians(a1.astype(np.float64))).astype(np.float32)
    arcs[..., 7] = np.sin(np.radians(a1.astype(np.float64))).astype(np.float32)
    arcs[..., 8] = width
    arcs[..., 9] = ((a1 - a0) % 360).astype(np.float32)
    return arcs


def arc_mask(H: int, W: int, arc: np.ndarray) -> np.ndarray:
    """Boolean (H, W) mask of one arc (float32 arithmetic, same operation order as the CUDA kernel)."""
    f32 = np.float32
    cx, cy, rx, ry, c0, s0, c1, s1, wd, sweep = (f32(v) for v in arc)
    yy, xx = np.mgrid[0:H, 0:W]
    dx = xx.astype(f32) - cx
    dy = yy.astype(f32) - cy
    u = dx / rx
    v = dy / ry
    outer = (u * u + v * v) <= f32(1)
    irx, iry = rx - wd, ry - wd
    if irx > 0 and iry > 0:
        ui, vi = dx / irx, dy / iry
        inner = (ui * ui + vi * vi) < f32(1)
    else:
        inner = np.zeros_like(outer)
    a = c0 * v - s0 * u       # sin(theta_p - theta_start)
    b = s1 * u - c1 * v       # sin(theta_end - theta_p)
    sector = ((a >= 0) & (b >= 0)) if sweep <= 180 else ~((a < 0) & (b < 0))
    return outer & ~inner & sector


def draw_spaghetti_u8(images_u8: np.ndarray, arcs: np.ndarray) -> np.ndarray:
    """images (n, H, W, 3) uint8, arcs (n, n_arcs, ARC_FIELDS) -> copy with the arcs painted black."""
    out = images_u8.copy()
    n, H, W, _ = out.shape
    for i in range(n):
        m = np.zeros((H, W), dtype=bool)
        for arc in arcs[i]:
            m |= arc_mask(H, W, arc)
        out[i][m] = 0
    return out
