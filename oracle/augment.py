"""ORACLE (test infrastructure only): CPU restatement of the reference's augmentation chain.

Reference call sites: /root/reference/argus/data.py:41-103 (Augmentation builds a kornia AugmentationSequential) and
data.py:213-225 (`/255`, then the chain on a (n_cams, 3, H, W) float image pair). Stages, in order (data.py:52-95):
[RandomErasing(p=.5, scale (0.02,0.1), ratio (2,3), value 0), RandomErasing(p=.5, scale (0.02,0.05), ratio (0.8,1.2),
value 1)] (flag random_erasing, default off) -> RandomPlanckianJitter("blackbody") p=.5 -> ColorJiggle(brightness
(0.8,1), contrast (0.5,1.2), saturation (0.25,1.2), hue (-0.1,0.1), same_on_batch=True, p=1) -> RandomGaussianBlur((5,5),
(3,8), p=.5) -> RandomMotionBlur(3, 35deg, 0.5, p=.7) -> RandomPlasmaShadow(roughness (0.1,0.4), intensity (-0.6,0),
quantity (0,0.5), p=1) -> [RandomSaltAndPepperNoise(p=.7)] (flag salt_and_pepper, default off).

The arithmetic lives in the un-vendored third-party dependency **kornia** (`kornia>=0.7.2`,
/root/reference/pyproject.toml:20; not installed, no network). The per-op semantics below restate kornia 0.7.x's
published algorithms (enhance.adjust_*, color.rgb_to_hsv/hsv_to_rgb, filters.gaussian/motion kernels, the blackbody
illuminant table, RectangleEraseGenerator, contrib.diamond_square: seed grid -> recursive diamond / square steps with
`(1 - scale) * neighbour mean + scale * U[0,1)`, scale multiplied by `roughness` per level, 4/3 border compensation of the
square step, crop of the 2^k + 1 grid to H x W, shadow where the field < shade_quantity; salt & pepper: one noise mask per
pixel shared by the channels). What cannot be reproduced, and is OUR frozen spec instead:
  * random numbers: kornia draws from torch's global RNG; here every parameter is a pure function of
    (seed, step, image index, field) through a splitmix64 hash, and every per-pixel draw (fractal, noise) a 32-bit hash
    of (image seed, level, y, x), identical in numpy and in the CUDA kernels;
  * details of kornia that could not be re-read offline (the 3x3 seed grid of diamond_square is taken as all-uniform,
    the square step's border handling as zero padding x 4/3).
No reference test checks an augmented pixel, so: **parity with kornia itself is unpinned**; what is
pinned is GPU == this oracle on identical parameters (tests/test_augment_gpu.py: discrete parts -- parameter tables,
erasing rectangles, shadow masks, noise masks, arcs -- bit for bit, float arithmetic to 1e-5), parameter ranges, and the
reference's only augmentation-related contract: same seed => same result (tests/test_train.py:69-77). The spaghetti arcs
ARE pinned to the reference's library: oracle/pil_arc.py reproduces Pillow's ImageDraw.arc pixel for pixel.
"""
from __future__ import annotations

import numpy as np

N_PARAMS = 40  # floats per image, layout shared with argus_b200/csrc/augment.cu (ARGUS_AUG_PARAMS)

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


def _erase_rect(u_area, u_ra, u_rb, u_pick, u_x, u_y, scale, ratio, H, W):
    """kornia RectangleEraseGenerator: area = U(scale) * H * W, aspect = h / w; -> (x, y, w, h) as float32 arrays."""
    area = _lerp(u_area, scale[0], scale[1]) * F(H * W)
    if ratio[0] < 1.0 and ratio[1] > 1.0:
        r1, r2 = _lerp(u_ra, ratio[0], 1.0), _lerp(u_rb, 1.0, ratio[1])
        aspect = np.where(np.rint(u_pick) != 0, r1, r2).astype(np.float32)
    else:
        aspect = _lerp(u_ra, ratio[0], ratio[1])
    h = np.rint(np.sqrt(area * aspect, dtype=np.float32))
    w = np.rint(np.sqrt(area / aspect, dtype=np.float32))
    h = np.minimum(np.maximum(h, F(1)), F(H)).astype(np.float32)
    w = np.minimum(np.maximum(w, F(1)), F(W)).astype(np.float32)
    x = np.floor(u_x * (F(W) - w + F(1))).astype(np.float32)
    y = np.floor(u_y * (F(H) - h + F(1))).astype(np.float32)
    return x, y, w, h


def sample_params(n_pairs: int, n_cams: int, seed: int, step: int, cfg=None, H: int = 256, W: int = 256) -> np.ndarray:
    """(n_pairs*n_cams, N_PARAMS) float32 parameter table. Image index = pair*n_cams + view; the colour-jiggle
    draws use the PAIR index (same_on_batch=True on the (n_cams,3,H,W) mini-batch, data.py:76,224).
    Layout: 0-1 planckian R / B gains; 2-6 jiggle (brightness add, contrast, saturation, hue [rad], order index or -1);
    7 gaussian sigma (0 = off); 8-16 motion kernel; 17-20 plasma (roughness, intensity (0 = off), quantity, seed);
    24-28 / 29-33 erasing rectangle 1 / 2 (on, x, y, w, h); 34-37 salt & pepper (on, amount, salt share, seed)."""
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
    # random erasing (data.py:52-64)
    if c["random_erasing"]:
        for e, (scale, ratio) in enumerate((((0.02, 0.1), (2.0, 3.0)), ((0.02, 0.05), (0.8, 1.2)))):
            f = 17 + 8 * e
            on = u(img, 16 + 8 * e) < np.float32(0.5)
            x, y, w, h = _erase_rect(u(img, f), u(img, f + 1), u(img, f + 2), u(img, f + 3), u(img, f + 4),
                                     u(img, f + 5), scale, ratio, H, W)
            base = 24 + 5 * e
            P[:, base] = on
            for k_, v in enumerate((x, y, w, h)):
                P[:, base + 1 + k_] = np.where(on, v, np.float32(0))
    # salt & pepper (data.py:94-95; kornia defaults amount (0.01, 0.06), salt_vs_pepper (0.4, 0.6))
    if c["salt_and_pepper"]:
        on = u(img, 32) < np.float32(0.7)
        P[:, 34] = on
        P[:, 35] = np.where(on, _lerp(u(img, 33), 0.01, 0.06), np.float32(0))
        P[:, 36] = np.where(on, _lerp(u(img, 34), 0.4, 0.6), np.float32(0))
    P[:, 37] = u(img, 35)
    return P


def _cfg(cfg):
    d = dict(brightness=(0.8, 1.0), contrast=(0.5, 1.2), saturation=(0.25, 1.2), hue=(-0.1, 0.1),
             color_jiggle=True, planckian_jitter=True, blur=True, motion_blur=True, plasma_shadow=True,
             random_erasing=False, salt_and_pepper=False)
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


def pixel_uniform(seed32: int, level: int, y, x) -> np.ndarray:
    """float32 uniform in [0,1) from a 32-bit hash of (image seed, level, y, x) -- csrc/augment.cu::pixel_uniform."""
    with np.errstate(over="ignore"):
        h = np.uint32(seed32 & 0xFFFFFFFF) ^ (np.uint32(level) * np.uint32(0x9E3779B9))
        h = (h ^ np.asarray(y, dtype=np.uint32)) * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(15))
        h = (h ^ np.asarray(x, dtype=np.uint32)) * np.uint32(0xC2B2AE35)
        h = h ^ (h >> np.uint32(13))
        h = h * np.uint32(0x27D4EB2F)
        h = h ^ (h >> np.uint32(16))
    return (h >> np.uint32(8)).astype(np.float32) * F(2.0 ** -24)


def _ceil_log2(v: int) -> int:
    l = 0
    while (1 << l) < v:
        l += 1
    return l


def plasma_field(H: int, W: int, roughness, seed32: int) -> np.ndarray:
    """kornia.contrib.diamond_square((1, 1, H, W), roughness) restated (see the module docstring): (H, W) float32."""
    lh, lw = _ceil_log2(H - 1), _ceil_log2(W - 1)
    depth = min(lh, lw) - 1
    sh, sw = (1 << (lh - depth)) + 1, (1 << (lw - depth)) + 1
    yy, xx = np.mgrid[0:sh, 0:sw]
    img = pixel_uniform(seed32, 0, yy, xx)
    rough = F(roughness)
    scale = F(1)
    comp = F(1.0 / 0.75)
    for level in range(1, depth + 1):
        scale = F(scale * rough)
        h, w = img.shape
        nh, nw = 2 * h - 1, 2 * w - 1
        new = np.zeros((nh, nw), dtype=np.float32)
        new[::2, ::2] = img
        yy, xx = np.mgrid[0:nh, 0:nw]
        rnd = pixel_uniform(seed32, level, yy, xx)
        one_minus = F(F(1) - scale)
        # diamond step: (odd, odd) = (1 - scale) * mean of the four diagonal parents + scale * u
        m = F(0.25) * (((img[:-1, :-1] + img[:-1, 1:]) + img[1:, :-1]) + img[1:, 1:])
        new[1::2, 1::2] = one_minus * m + scale * rnd[1::2, 1::2]
        # square step: positions with exactly one odd coordinate; zero padding, border rows / columns x 4/3
        p = np.pad(new, 1)
        up, down, left, right = p[:-2, 1:-1], p[2:, 1:-1], p[1:-1, :-2], p[1:-1, 2:]
        reg = F(0.25) * (((up + left) + right) + down)
        border = np.zeros((nh, nw), dtype=bool)
        border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = True
        reg = np.where(border, reg * comp, reg).astype(np.float32)
        sq = ((yy ^ xx) & 1) == 1
        new = np.where(sq, one_minus * reg + scale * rnd, new).astype(np.float32)
        img = new
    return img[:H, :W]


def plasma_shadow_mask(H: int, W: int, p: np.ndarray) -> np.ndarray:
    """Boolean (H, W): where RandomPlasmaShadow darkens the image (field < shade_quantity)."""
    if p[18] == 0:
        return np.zeros((H, W), dtype=bool)
    return plasma_field(H, W, p[17], int(round(float(p[20]) * (1 << 24)))) < p[19]


def salt_pepper_masks(H: int, W: int, p: np.ndarray):
    """(salt, pepper) boolean (H, W) masks of RandomSaltAndPepperNoise; one draw per pixel, shared by the channels."""
    if p[34] == 0:
        z = np.zeros((H, W), dtype=bool)
        return z, z
    seed32 = int(round(float(p[37]) * (1 << 24)))
    yy, xx = np.mgrid[0:H, 0:W]
    noise = pixel_uniform(seed32, 100, yy, xx) < p[35]
    salt = pixel_uniform(seed32, 101, yy, xx) < p[36]
    return noise & salt, noise & ~salt


def erase(x: np.ndarray, p: np.ndarray) -> np.ndarray:
    """The two RandomErasing rectangles (value 0, then value 1) on a (3, H, W) image."""
    x = x.copy()
    for base, value in ((24, 0.0), (29, 1.0)):
        if p[base] != 0:
            ex, ey, ew, eh = (int(v) for v in p[base + 1:base + 5])
            x[:, ey:ey + eh, ex:ex + ew] = F(value)
    return x


def augment_image(u8_hwc: np.ndarray, p: np.ndarray, arc_mask: np.ndarray | None = None) -> np.ndarray:
    """One image: uint8 (H,W,3) -> float32 (3,H,W) in [0,1], parameters p (N_PARAMS,); arc_mask: optional boolean
    (H, W) spaghetti mask painted black first (data.py:212-215 draws the arcs on the decoded image)."""
    u8 = u8_hwc
    if arc_mask is not None:
        u8 = u8.copy()
        u8[arc_mask] = 0
    x = (u8.astype(np.float32) * F(1.0 / 255.0)).transpose(2, 0, 1)
    x = erase(x, p)
    x = color_ops(x, p)
    x = gaussian_blur(x, float(p[7]))
    x = motion_blur(x, p[8:17])
    H, W = x.shape[1:]
    x = x + np.where(plasma_shadow_mask(H, W, p), p[18], F(0))[None]
    x = np.clip(x, F(0), F(1)).astype(np.float32)
    salt, pepper = salt_pepper_masks(H, W, p)
    x[:, salt] = F(1)
    x[:, pepper] = F(0)
    return x


def augment_batch_u8(images_u8: np.ndarray, seed: int = 0, step: int = 0, cfg=None, params=None) -> np.ndarray:
    """(B, n_cams, H, W, 3) uint8 -> (B, n_cams, 3, H, W) float32 (what Dataset.__getitem__ hands to the model after
    `reshape(-1, H, W)`, data.py:224-227)."""
    B, n_cams, H, W, _ = images_u8.shape
    if params is None:
        params = sample_params(B, n_cams, seed, step, cfg, H=H, W=W)
    out = np.empty((B, n_cams, 3, H, W), dtype=np.float32)
    for b in range(B):
        for v in range(n_cams):
            out[b, v] = augment_image(images_u8[b, v], params[b * n_cams + v])
    return out


# ------------------------------------------------------------------------------------------------------------------
# spaghetti arcs (reference: argus/utils.py:252-275 `draw_spaghetti`, applied at argus/data.py:212-215 to the decoded
# image BEFORE the kornia chain). The reference draws with PIL's ImageDraw.arc; oracle/pil_arc.py restates Pillow's
# rasteriser and is pinned to the real library pixel for pixel (tests/test_oracle_augment.py).
# Sampling follows the reference: x0 ~ U{0..W-1}, y0 ~ U{0..H-1}, x1 ~ U{x0..W-1}, y1 ~ U{y0..H-1},
# start, end ~ U{0..359}, width = int(U(1, 5)); one draw per (seed, step, image, arc, field) through the same hash.
# ------------------------------------------------------------------------------------------------------------------
ARC_FIELDS = 8   # x0, y0, x1, y1, start, end, width, 0
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
    for k, v in enumerate((x0, y0, x1, y1, a0, a1)):
        arcs[..., k] = v.astype(np.float32)
    arcs[..., 6] = width
    return arcs


def arc_mask(H: int, W: int, arc: np.ndarray) -> np.ndarray:
    """Boolean (H, W) mask of one arc row of the table: exactly the pixels Pillow's ImageDraw.arc paints."""
    from .pil_arc import arc_mask as pil_arc_mask

    x0, y0, x1, y1, a0, a1, wd = (int(v) for v in arc[:7])
    return pil_arc_mask(H, W, (x0, y0, x1, y1), a0, a1, wd)


def spaghetti_mask(H: int, W: int, arcs: np.ndarray) -> np.ndarray:
    """arcs (n, n_arcs, ARC_FIELDS) -> boolean (n, H, W)."""
    out = np.zeros((arcs.shape[0], H, W), dtype=bool)
    for i in range(arcs.shape[0]):
        for arc in arcs[i]:
            out[i] |= arc_mask(H, W, arc)
    return out


def draw_spaghetti_u8(images_u8: np.ndarray, arcs: np.ndarray) -> np.ndarray:
    """images (n, H, W, 3) uint8, arcs (n, n_arcs, ARC_FIELDS) -> copy with the arcs painted black."""
    out = images_u8.copy()
    n, H, W, _ = out.shape
    out[spaghetti_mask(H, W, arcs)] = 0
    return out
