"""GPU parity of the augmentation kernels against the numpy oracle (oracle/augment.py, oracle/pil_arc.py) on identical
parameters: parameter tables, plasma shadow masks, erasing / noise and the spaghetti arcs bit for bit (integer / discrete
work), the float image arithmetic to 1e-5 (fast-math divisions on the GPU), plus the staging layout."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_u8(B, n_cams, H, W, seed):
    rng = np.random.default_rng(seed)
    # structured + noise so that blur / hue / saturation all matter
    yy, xx = np.mgrid[0:H, 0:W]
    base = (np.sin(xx / 9.0)[None, None, :, :, None] * 60 + np.cos(yy / 13.0)[None, None, :, :, None] * 50 + 120)
    img = base + rng.normal(0, 25, (B, n_cams, H, W, 3)) + rng.uniform(-40, 40, (B, n_cams, 1, 1, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def unpack_bits(mask: torch.Tensor, W: int) -> np.ndarray:
    """int32 [n, H, W // 32] -> boolean (n, H, W)."""
    m = mask.cpu().numpy().view(np.uint32)
    bits = (m[..., None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return bits.reshape(m.shape[0], m.shape[1], W).astype(bool)


ALL_ON = dict(random_erasing=True, salt_and_pepper=True)


def test_param_table_matches_oracle(cuda_device):
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as oracle_aug

    for kw, H, W in ((dict(), 256, 256), (ALL_ON, 256, 256), (ALL_ON, 128, 64)):
        cfg = AugmentationConfig(**kw)
        aug = Augmentation(cfg, train=True, seed=1234)
        for step in (0, 7):
            got = aug.sample_params(64, 2, cuda_device, step=step, H=H, W=W).cpu().numpy()
            want = oracle_aug.sample_params(64, 2, seed=1234, step=step, cfg=cfg, H=H, W=W)
            # identical hash, identical fp32 arithmetic except libm ulps in cos/sin/normalisation of the motion kernel
            assert np.array_equal(got[:, :8], want[:, :8])
            assert np.allclose(got[:, 8:17], want[:, 8:17], rtol=0, atol=2e-7)
            assert np.array_equal(got[:, 17:], want[:, 17:])          # plasma, erasing rectangles, salt & pepper


@pytest.mark.parametrize("H,W", [(64, 64), (256, 256), (128, 64)])
@pytest.mark.parametrize("kw", [dict(), ALL_ON])
def test_augment_matches_oracle(cuda_device, H, W, kw):
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as oracle_aug

    B, n_cams = 4, 2
    u8 = make_u8(B, n_cams, H, W, seed=H + W)
    cfg = AugmentationConfig(**kw)
    aug = Augmentation(cfg, train=True, seed=99)
    params = aug.sample_params(B, n_cams, cuda_device, step=3, H=H, W=W)
    got = aug.augment_batch(torch.from_numpy(u8).to(cuda_device), params=params).cpu().numpy()
    P = params.cpu().numpy()
    # the discrete part first: the shadow mask the GPU built (left in the workspace) == the oracle's diamond-square mask
    ws = aug.workspace(B * n_cams, H, W, cuda_device)
    gpu_mask = unpack_bits(ws, W)
    for i in range(B * n_cams):
        assert np.array_equal(gpu_mask[i], oracle_aug.plasma_shadow_mask(H, W, P[i])), i
    want = oracle_aug.augment_batch_u8(u8, params=P).reshape(B, 3 * n_cams, H, W)
    diff = np.abs(got - want)
    # fp32 arithmetic on both sides (fast-math division / exp on the GPU); isolated pixels may flip an HSV sector at ties
    assert np.mean(diff > 1e-5) < 2e-3, (np.mean(diff > 1e-5), diff.max())
    assert np.mean(diff) < 1e-5
    assert got.min() >= 0.0 and got.max() <= 1.0
    if kw:
        # erasing rectangles and noise pixels are exact
        for i in range(B * n_cams):
            b, v = divmod(i, n_cams)
            salt, pepper = oracle_aug.salt_pepper_masks(H, W, P[i])
            g = got[b, 3 * v:3 * v + 3]
            assert (g[:, salt] == 1).all() and (g[:, pepper] == 0).all()


def test_float_input_and_identity(cuda_device):
    """Reference call convention: float (n_cams,3,H,W) in -> same shape out; identity when train=False
    (reference data.py:99-103)."""
    from argus_b200.data import Augmentation, AugmentationConfig

    x = torch.rand(2, 3, 64, 64, device=cuda_device)
    assert torch.equal(Augmentation(AugmentationConfig(), train=False)(x), x)
    aug = Augmentation(AugmentationConfig(), train=True, seed=5)
    y = aug(x)
    assert y.shape == x.shape and not torch.equal(y, x)
    # same seed -> same draw sequence (reference tests/test_train.py:69-77 relies on this)
    aug2 = Augmentation(AugmentationConfig(), train=True, seed=5)
    assert torch.equal(aug2(x), y)
    assert not torch.equal(aug2(x), y)  # the step counter advanced
    off = Augmentation(AugmentationConfig(color_jiggle=False, planckian_jitter=False, blur=False, motion_blur=False,
                                          plasma_shadow=False), train=True)
    assert torch.equal(off(x), x)
    # the default-off stages of the reference construct and run (they used to raise NotImplementedError)
    full = Augmentation(AugmentationConfig(random_erasing=True, salt_and_pepper=True), train=True, seed=5)
    z = full(x)
    assert z.shape == x.shape and torch.isfinite(z).all() and z.min() >= 0 and z.max() <= 1


def test_staged_input_equals_explicit_path(cuda_device):
    """Fused augmentation+staging (uint8 -> bf16 space-to-depth inside the model) must feed the network the same
    pixels as augment_batch() followed by the fp32 NCHW entry point, with and without the spaghetti mask."""
    from argus_b200.data import Augmentation, AugmentationConfig
    from argus_b200.models import NCameraCNN

    torch.manual_seed(0)
    model = NCameraCNN().to(cuda_device).eval()
    B, n_cams, H, W = 2, 2, 128, 128
    u8 = torch.from_numpy(make_u8(B, n_cams, H, W, seed=1)).to(cuda_device)
    aug = Augmentation(AugmentationConfig(), train=True, seed=7, gpu_spaghetti=True)
    params = aug.sample_params(B, n_cams, cuda_device, step=0, H=H, W=W)
    arcs = aug.arc_params(B * n_cams, H, W, cuda_device, step=0)
    mask = aug.arc_mask(B * n_cams, H, W, cuda_device, arcs=arcs).clone()
    with torch.no_grad():
        y_fused = model._forward_impl(u8, False, aug_params=params.clone(), augment=True, arc_mask=mask)
        x = aug.augment_batch(u8, params=params.clone(), arc_mask=mask)
        y_explicit = model(x)
        y_two_step = model(aug.augment_batch(aug.spaghetti_batch(u8, step=0), params=params.clone()))
        y_noarc = model._forward_impl(u8, False, aug_params=params.clone(), augment=True)
        y_plain = model(u8)  # no augmentation: only /255 and packing
        y_plain_f32 = model(u8.permute(0, 1, 4, 2, 3).reshape(B, 6, H, W).float() / 255.0)
        y_arcs_only = model._forward_impl(u8, False, arc_mask=mask)
        y_arcs_only_ref = model(aug.spaghetti_batch(u8, step=0))
    assert torch.allclose(y_fused, y_explicit, rtol=1e-3, atol=1e-5)
    assert torch.allclose(y_fused, y_two_step, rtol=1e-3, atol=1e-5)      # mask fused into the kernel == drawn first
    assert torch.allclose(y_plain, y_plain_f32, rtol=1e-3, atol=1e-5)
    assert torch.allclose(y_arcs_only, y_arcs_only_ref, rtol=1e-3, atol=1e-5)
    assert not torch.allclose(y_fused, y_plain, rtol=1e-3, atol=1e-5)
    assert not torch.allclose(y_fused, y_noarc, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("H,W", [(128, 128), (256, 256), (64, 160)])
def test_spaghetti_gpu_is_pillow_exact(cuda_device, H, W):
    """GPU arc rasteriser == the oracle (== Pillow's ImageDraw.arc, tests/test_oracle_augment.py) bit for bit on the same
    (seed, step): parameter table, bit mask and painted pixels. 256 x 256 is the reference's image size."""
    from argus_b200 import _lib
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as A

    B, n_cams = 3, 2
    g = torch.Generator().manual_seed(2)
    images = torch.randint(1, 256, (B, n_cams, H, W, 3), dtype=torch.uint8, generator=g).to(cuda_device)
    aug = Augmentation(AugmentationConfig(), train=True, seed=11, gpu_spaghetti=True)
    arcs = aug.arc_params(B * n_cams, H, W, cuda_device, step=7)
    want_arcs = A.spaghetti_params(B * n_cams, 10, H, W, seed=aug.seed, step=7)
    assert np.array_equal(arcs.cpu().numpy(), want_arcs)
    want_mask = A.spaghetti_mask(H, W, want_arcs)
    got_mask = unpack_bits(aug.arc_mask(B * n_cams, H, W, cuda_device, arcs=arcs), W)
    assert np.array_equal(got_mask, want_mask), int((got_mask ^ want_mask).sum())
    out = aug.spaghetti_batch(images, step=7)
    want = A.draw_spaghetti_u8(images.cpu().numpy().reshape(B * n_cams, H, W, 3), want_arcs).reshape(B, n_cams, H, W, 3)
    got = out.cpu().numpy()
    assert np.array_equal(got, want)
    assert 0.005 < (got == 0).all(-1).mean() < 0.3          # arcs were drawn, and not everywhere
    assert torch.equal(images.cpu(), torch.randint(1, 256, (B, n_cams, H, W, 3), dtype=torch.uint8,
                                                   generator=torch.Generator().manual_seed(2)))   # input untouched


def test_spaghetti_gpu_against_real_pillow(cuda_device):
    """End to end against the library the reference calls: GPU mask of hand-picked and random arcs == PIL's pixels."""
    from PIL import Image, ImageDraw

    from argus_b200 import _lib

    H = W = 256
    rng = np.random.default_rng(0)
    n = 64
    arcs = np.zeros((n, 1, 8), dtype=np.float32)
    for i in range(n):
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0, W)), int(rng.integers(y0, H))
        arcs[i, 0, :7] = [x0, y0, x1, y1, int(rng.integers(0, 360)), int(rng.integers(0, 360)), int(rng.uniform(1, 5))]
    arcs[0, 0, :7] = [0, 0, 255, 255, 0, 360 - 1, 4]
    arcs[1, 0, :7] = [10, 40, 250, 60, 350, 10, 3]       # flat, wraps through 0 degrees
    arcs[2, 0, :7] = [100, 5, 104, 250, 89, 271, 2]      # tall (transposed clip tree), near the axes
    arcs[3, 0, :7] = [7, 7, 7, 7, 10, 200, 1]            # degenerate bbox
    t = torch.from_numpy(arcs).to(cuda_device)
    mask = torch.empty((n, H, W // 32), dtype=torch.int32, device=cuda_device)
    _lib.call("argus_spaghetti_mask", t, mask, n, 1, H, W, _lib.stream_ptr())
    got = unpack_bits(mask, W)
    for i in range(n):
        x0, y0, x1, y1, a0, a1, wd = (int(v) for v in arcs[i, 0, :7])
        img = Image.new("L", (W, H), 255)
        ImageDraw.Draw(img).arc((x0, y0, x1, y1), a0, a1, fill=0, width=wd)
        assert np.array_equal(got[i], np.array(img) == 0), (i, arcs[i, 0])
