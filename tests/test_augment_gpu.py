"""GPU parity of the fused augmentation kernel against the numpy oracle (oracle/augment.py) on identical
parameters, plus parameter-table parity (the RNG is a shared splitmix64 hash) and the staging layout."""
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


def test_param_table_matches_oracle(cuda_device):
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as oracle_aug

    aug = Augmentation(AugmentationConfig(), train=True, seed=1234)
    for step in (0, 7):
        got = aug.sample_params(64, 2, cuda_device, step=step).cpu().numpy()
        want = oracle_aug.sample_params(64, 2, seed=1234, step=step)
        # identical hash, identical fp32 arithmetic except libm ulps in cos/sin/normalisation of the motion kernel
        assert np.array_equal(got[:, :8], want[:, :8])
        assert np.allclose(got[:, 8:17], want[:, 8:17], rtol=0, atol=2e-7)
        assert np.array_equal(got[:, 17:21], want[:, 17:21])


@pytest.mark.parametrize("H,W", [(64, 64), (256, 256), (128, 64)])
def test_augment_matches_oracle(cuda_device, H, W):
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as oracle_aug

    B, n_cams = 4, 2
    u8 = make_u8(B, n_cams, H, W, seed=H + W)
    aug = Augmentation(AugmentationConfig(), train=True, seed=99)
    params = aug.sample_params(B, n_cams, cuda_device, step=3)
    got = aug.augment_batch(torch.from_numpy(u8).to(cuda_device), params=params).cpu().numpy()
    want = oracle_aug.augment_batch_u8(u8, params=params.cpu().numpy()[:, :]).reshape(B, 3 * n_cams, H, W)
    diff = np.abs(got - want)
    # fp32 arithmetic on both sides; isolated pixels may flip the plasma threshold or an HSV sector at 1-ulp ties
    assert np.mean(diff > 1e-5) < 2e-3, (np.mean(diff > 1e-5), diff.max())
    assert np.mean(diff) < 1e-5
    assert got.min() >= 0.0 and got.max() <= 1.0


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


def test_staged_input_equals_explicit_path(cuda_device):
    """Fused augmentation+staging (uint8 -> bf16 space-to-depth inside the model arena) must feed the network the same
    pixels as augment_batch() followed by the fp32 NCHW entry point."""
    from argus_b200.data import Augmentation, AugmentationConfig
    from argus_b200.models import NCameraCNN

    torch.manual_seed(0)
    model = NCameraCNN().to(cuda_device).eval()
    B, n_cams, H, W = 2, 2, 128, 128
    u8 = torch.from_numpy(make_u8(B, n_cams, H, W, seed=1)).to(cuda_device)
    aug = Augmentation(AugmentationConfig(), train=True, seed=7)
    params = aug.sample_params(B, n_cams, cuda_device, step=0)
    with torch.no_grad():
        y_fused = model._forward_impl(u8, False, aug_params=params.clone(), augment=True)
        x = aug.augment_batch(u8, params=params.clone())
        y_explicit = model(x)
        y_plain = model(u8)  # no augmentation: only /255 and packing
        y_plain_f32 = model(u8.permute(0, 1, 4, 2, 3).reshape(B, 6, H, W).float() / 255.0)
    assert torch.allclose(y_fused, y_explicit, rtol=1e-3, atol=1e-5)
    assert torch.allclose(y_plain, y_plain_f32, rtol=1e-3, atol=1e-5)
    assert not torch.allclose(y_fused, y_plain, rtol=1e-3, atol=1e-5)


def test_spaghetti_gpu_matches_oracle(cuda_device):
    """GPU arc rasteriser == numpy oracle bit for bit on the same (seed, step): parameter table and painted pixels."""
    import numpy as np

    from argus_b200 import _lib
    from argus_b200.data import Augmentation, AugmentationConfig
    from oracle import augment as A

    B, n_cams, H, W = 3, 2, 128, 128
    g = torch.Generator().manual_seed(2)
    images = torch.randint(1, 256, (B, n_cams, H, W, 3), dtype=torch.uint8, generator=g).to(cuda_device)
    aug = Augmentation(AugmentationConfig(), train=True, seed=11, gpu_spaghetti=True)
    out = aug.spaghetti_batch(images, step=7)
    arcs = torch.empty(B * n_cams, 10, 10, device=cuda_device)
    import ctypes
    _lib.check(_lib.load().argus_spaghetti_sample_params(_lib.ptr(arcs), ctypes.c_int(B * n_cams), ctypes.c_int(10),
                                                         ctypes.c_int(H), ctypes.c_int(W), ctypes.c_uint64(aug.seed),
                                                         ctypes.c_uint64(7), _lib.stream_ptr()))
    want_arcs = A.spaghetti_params(B * n_cams, 10, H, W, seed=aug.seed, step=7)
    assert np.array_equal(arcs.cpu().numpy(), want_arcs)
    want = A.draw_spaghetti_u8(images.cpu().numpy().reshape(B * n_cams, H, W, 3), want_arcs).reshape(B, n_cams, H, W, 3)
    got = out.cpu().numpy()
    assert np.array_equal(got, want)
    assert 0.01 < (got == 0).all(-1).mean() < 0.3          # arcs were drawn, and not everywhere
    assert torch.equal(images.cpu(), torch.randint(1, 256, (B, n_cams, H, W, 3), dtype=torch.uint8,
                                                   generator=torch.Generator().manual_seed(2)))   # input untouched
