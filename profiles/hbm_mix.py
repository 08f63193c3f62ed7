"""HBM bandwidth by read/write mix on this GPU (the roofline denominators of the memory-bound kernels depend on it:
MEASURED_PEAKS.json's hbm_gbs is a 1:1 copy). Probes: pure write (zero_), 1:1 copy (copy_), 2:1 read/write and pure
read (the library's own bn_bwd_apply / bn_bwd_reduce ring kernels). 1 GiB bf16 tensors, CUDA events, best of 10."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from argus_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
rows, C = 1 << 22, 128          # 1 GiB per bf16 tensor
a = torch.randn(rows, C, device=dev).bfloat16()
b = torch.empty_like(a)
c = torch.empty_like(a)
sc = torch.ones(C, device=dev); z = torch.zeros(C, device=dev)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
nbytes = a.numel() * 2


def best(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


res = {}
res["write only (zero_)"] = nbytes / best(lambda: b.zero_()) / 1e6
res["copy 1R:1W (copy_)"] = 2 * nbytes / best(lambda: b.copy_(a)) / 1e6
lib = _lib.load()
scratch_apply = lambda: _lib.check(lib.argus_bn_backward(_lib.ptr(a), _lib.ptr(b), None, _lib.ptr(sc), _lib.ptr(z), _lib.ptr(z),  # noqa: E731
                                                         _lib.ptr(sc), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(c), _lib.c_int64(rows),
                                                         _lib.c_int(C), _lib.c_int(0), _lib.stream_ptr()))
# argus_bn_backward = reduce (2 reads) + finalize + apply (2 reads, 1 write): 5 tensor passes, 4R:1W overall
res["bn_backward 4R:1W (reduce + apply rings)"] = 5 * nbytes / best(scratch_apply) / 1e6
for k, v in res.items():
    print(f"{k:44s} {v:8.1f} GB/s")
