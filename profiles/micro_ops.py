"""Micro-benchmark / ncu target: single hot-path ops at the bench shapes (B=256 pairs -> 512 images), through the C ABI.
Usage: python profiles/micro_ops.py [reps]   (prints CUDA-event times; run under ncu with -k regex:... for captures)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from argus_b200 import _lib  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
N = 512


def timeit(name, fn, bytes_=0.0, flops=0.0):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    extra = ""
    if bytes_:
        extra += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TF/s"
    print(f"{name:40s} {ms:8.4f} ms{extra}", flush=True)


def bf(*shape):
    return torch.randn(*shape, device=dev).bfloat16()


# layer1 3x3 convolution 64 -> 64 at 64x64: forward and dgrad
x = bf(N, 64, 64, 64); w = bf(64, 3, 3, 64); y = torch.empty_like(x)
fl = 2.0 * N * 64 * 64 * 64 * 576
timeit("conv3x3 64->64 @64 fwd", lambda: _lib.call("argus_conv2d_forward", x, w, y, N, 64, 64, 64, 64, 3, 1, 0, None, None, None, 0, None, 0, _lib.stream_ptr()), flops=fl)
timeit("conv3x3 64->64 @64 dgrad", lambda: _lib.call("argus_conv2d_dgrad", x, w, y, N, 64, 64, 64, 64, 3, 1, None, _lib.stream_ptr()), flops=fl)
# layer1 1x1 convolution 64 -> 256 at 64x64 forward (HBM-bound, wide output)
w2 = bf(256, 64); y2 = bf(N, 64, 64, 256)
timeit("conv1x1 64->256 @64 fwd", lambda: _lib.call("argus_conv2d_forward", x, w2, y2, N, 64, 64, 64, 256, 1, 1, 0, None, None, None, 0, None, 0, _lib.stream_ptr()),
       bytes_=x.numel() * 2 + y2.numel() * 2, flops=2.0 * N * 4096 * 64 * 256)
# stem max pooling (fused BN + ReLU) forward / backward at 128x128x64
xs = bf(N, 128, 128, 64); sc = torch.rand(64, device=dev) + 0.5; sh = torch.randn(64, device=dev)
yp = torch.empty(N, 64, 64, 64, device=dev, dtype=torch.bfloat16); idx = torch.empty(N, 64, 64, 64, device=dev, dtype=torch.uint8)
timeit("maxpool fwd 128x128x64", lambda: _lib.call("argus_maxpool_forward", xs, sc, sh, yp, idx, N, 128, 128, 64, _lib.stream_ptr()),
       bytes_=xs.numel() * 2 + yp.numel() * 3)
dyp = bf(N, 64, 64, 64); dxs = torch.empty_like(xs)
timeit("maxpool bwd 128x128x64", lambda: _lib.call("argus_maxpool_backward", dyp, idx, dxs, N, 128, 128, 64, _lib.stream_ptr()),
       bytes_=xs.numel() * 2 + yp.numel() * 3)
