"""Per-kernel parity of the fp32 parity mode (through the C ABI) against torch float64 on the same inputs.
Tolerance: 1e-5 relative (|ours - truth|_2 / |truth|_2) -- fp32 arithmetic with fp32 FMA accumulation over K <= 4608
and fp64 reductions; the north star's fp32 tolerance is 1e-4."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from argus_b200 import _lib
from gpu_util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride", [
    (2, 32, 32, 3, 64, 7, 2),       # stem
    (2, 16, 16, 64, 64, 3, 1),
    (3, 16, 8, 128, 128, 3, 2),
    (2, 8, 8, 256, 1024, 1, 1),
    (2, 16, 16, 256, 512, 1, 2),    # downsample
    (5, 1, 1, 2048, 1024, 1, 1),    # fc
    (1, 9, 7, 16, 24, 3, 1),        # ragged sizes (tile tails in M, N and K)
])
def test_fp32_conv(cuda_device, N, H, W, Cin, Cout, k, stride):
    g = torch.Generator().manual_seed(N * H + Cin + k)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda_device)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(cuda_device)
    b = torch.randn(Cout, generator=g).to(cuda_device)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y64 = F.conv2d(x64, w64, b.double(), stride=stride, padding=k // 2)
    dy = torch.randn(y64.shape, generator=torch.Generator().manual_seed(7)).to(cuda_device)
    y64.backward(dy.double())
    Ho, Wo = y64.shape[2], y64.shape[3]
    xh, dyh = nhwc(x), nhwc(dy)
    y = torch.empty(N, Ho, Wo, Cout, device=cuda_device)
    _lib.call("argus_fp32_conv2d_forward", xh, w, b, y, N, H, W, Cin, Cout, k, stride, _lib.stream_ptr())
    assert rel(y, nhwc(y64.detach())) < TOL
    dx = torch.full((N, H, W, Cin), float("nan"), device=cuda_device)
    _lib.call("argus_fp32_conv2d_dgrad", dyh, w, dx, N, H, W, Cin, Cout, k, stride, _lib.stream_ptr())
    assert rel(dx, nhwc(x64.grad)) < TOL
    dw = torch.ones_like(w)        # accumulates: start from a known value
    _lib.call("argus_fp32_conv2d_wgrad", dyh, xh, dw, N, H, W, Cin, Cout, k, stride, _lib.stream_ptr())
    assert rel(dw - 1, w64.grad) < TOL * 3     # the "+1" start value costs a few ulps of the (small) gradient
    dw2 = torch.ones_like(w)
    _lib.call("argus_fp32_conv2d_wgrad", dyh, xh, dw2, N, H, W, Cin, Cout, k, stride, _lib.stream_ptr())
    assert torch.equal(dw, dw2)    # deterministic split-K


@pytest.mark.parametrize("rows,C,mode", [(4096, 64, "relu"), (1000, 256, "plain"), (513, 2048, "residual"),
                                          (777, 512, "downsample"), (2048, 128, "relu")])
def test_fp32_batch_norm(cuda_device, rows, C, mode):
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 2 + torch.randn(C, generator=g)).to(cuda_device)
    gamma = (torch.rand(C, generator=g) + 0.5).to(cuda_device)
    beta = torch.randn(C, generator=g).to(cuda_device)
    res = torch.randn(rows, C, generator=g).to(cuda_device)
    rs, rb = (torch.rand(C, generator=g) + 0.5).to(cuda_device), torch.randn(C, generator=g).to(cuda_device)
    dy = torch.randn(rows, C, generator=g).to(cuda_device)
    rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    scale, shift, mean, invstd = (torch.empty(C, device=cuda_device) for _ in range(4))
    lib = _lib.load()
    _lib.check(lib.argus_fp32_bn_train(_lib.ptr(x), ctypes.c_int64(rows), ctypes.c_int(C), _lib.ptr(gamma), _lib.ptr(beta),
                                       _lib.ptr(rm), _lib.ptr(rv), ctypes.c_float(0.1), ctypes.c_float(1e-5), _lib.ptr(scale),
                                       _lib.ptr(shift), _lib.ptr(mean), _lib.ptr(invstd), _lib.stream_ptr()))
    x64 = x.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    m64, v64 = x64.mean(0), x64.var(0, unbiased=False)
    y64 = (x64 - m64) / (v64 + 1e-5).sqrt() * g64 + b64
    assert rel(mean, m64.detach()) < TOL and rel(invstd, (v64.detach() + 1e-5).rsqrt()) < TOL
    assert rel(rm, 0.1 * m64.detach()) < TOL and rel(rv, 0.9 + 0.1 * x.double().var(0, unbiased=True)) < TOL
    if mode == "residual":
        y64 = y64 + res.double()
    elif mode == "downsample":
        y64 = y64 + res.double() * rs.double() + rb.double()
    relu = mode != "plain"
    out64 = y64.relu() if relu else y64
    out = torch.empty_like(x)
    use_res = mode in ("residual", "downsample")
    _lib.check(lib.argus_fp32_bn_apply(_lib.ptr(x), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(res if use_res else None),
                                       _lib.ptr(rs if mode == "downsample" else None), _lib.ptr(rb if mode == "downsample" else None),
                                       ctypes.c_int(int(relu)), _lib.ptr(out), ctypes.c_int64(rows), ctypes.c_int(C),
                                       _lib.stream_ptr()))
    assert rel(out, out64.detach()) < TOL
    # backward with the mask of OUR output (elements within rounding of zero must not flip between the two sides)
    mask = (out > 0).double() if relu else torch.ones_like(out64)
    (y64 * mask).backward(dy.double())
    dgamma, dbeta = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
    dx, gout = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.argus_fp32_bn_backward(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(out if relu else None), _lib.ptr(scale),
                                          _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dx),
                                          _lib.ptr(gout), ctypes.c_int64(rows), ctypes.c_int(C), _lib.stream_ptr()))
    assert rel(dx, x64.grad) < TOL
    assert rel(dgamma, g64.grad) < TOL and rel(dbeta, b64.grad) < TOL
    assert torch.equal(gout, dy * mask.float())


def test_fp32_maxpool(cuda_device):
    N, H, W, C = 2, 32, 16, 64
    x = torch.randn(N, C, H, W, device=cuda_device).relu()
    xr = x.double().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    y = torch.empty(N, H // 2, W // 2, C, device=cuda_device)
    idx = torch.empty(N, H // 2, W // 2, C, device=cuda_device, dtype=torch.uint8)
    _lib.call("argus_fp32_maxpool_forward", nhwc(x), y, idx, N, H, W, C, _lib.stream_ptr())
    assert torch.equal(y, nhwc(ref.detach().float()))
    dy = torch.randn(N, C, H // 2, W // 2, device=cuda_device)
    ref.backward(dy.double())
    dx = torch.empty(N, H, W, C, device=cuda_device)
    _lib.call("argus_fp32_maxpool_backward", nhwc(dy), idx, dx, N, H, W, C, _lib.stream_ptr())
    # ties (many exact zeros after ReLU) may route to a different, equally valid tap: compare per-window sums
    assert torch.allclose(dx.sum((1, 2)).double(), nhwc(xr.grad).sum((1, 2)), rtol=1e-5, atol=1e-5)
    pos = nhwc(x) > 0      # strictly positive maxima are unique almost surely
    assert rel(dx[pos], nhwc(xr.grad)[pos].float()) < TOL
