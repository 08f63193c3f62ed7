"""Experiment: does running producer -> consumer pairs chunk by chunk keep the intermediate in the 126 MB L2?
Pair = bn_apply (residual + ReLU, writes `out`) followed by the next block's 1x1 convolution reading `out`
(layer1 shapes: 2097152 pixels x 256 channels -> 64). Full-tensor launches vs `chunks` sub-batches, both replayed from a
CUDA graph so that host launch overhead does not matter."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from argus_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
N, H, W, C, Cm = 504, 64, 64, 256, 64      # 504 = 28 chunks of 18 images
rows = N * H * W
raw3 = torch.randn(rows, C, device=dev).bfloat16()
xin = torch.randn(rows, C, device=dev).bfloat16()
out = torch.empty_like(raw3)
bits = torch.empty(rows, C // 8, device=dev, dtype=torch.uint8)
sc = torch.rand(C, device=dev) + 0.5
sh = torch.randn(C, device=dev)
w = (torch.randn(Cm, C, device=dev) / 16).bfloat16()
raw1 = torch.empty(rows, Cm, device=dev, dtype=torch.bfloat16)
lib = _lib.load()


def pair(n_img, off_img):
    r0 = off_img * H * W
    r = n_img * H * W
    _lib.call("argus_bn_apply_bits", raw3[r0:r0 + r], sc, sh, xin[r0:r0 + r], None, None, 1, out[r0:r0 + r], bits[r0:r0 + r],
              _lib.c_int64(r), C, _lib.stream_ptr())
    _lib.call("argus_conv2d_forward", out[r0:r0 + r], w, raw1[r0:r0 + r], n_img, H, W, C, Cm, 1, 1, 0, None, None, None, 0,
              None, 0, _lib.stream_ptr())


def bench(chunk_imgs, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for off in range(0, N, chunk_imgs):
            pair(chunk_imgs, off)          # warm-up (also configures kernels outside capture)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for off in range(0, N, chunk_imgs):
                pair(chunk_imgs, off)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for chunk in (504, 252, 72, 36, 18, 9):
    ms = bench(chunk)
    print(f"chunk {chunk:4d} images ({N // chunk:3d} chunks, {chunk * H * W * C * 2 / 1e6:7.1f} MB intermediate): {ms:7.4f} ms", flush=True)
