"""Per-layer kernel times of the eval forward (BASELINE.json configs[2]): python profiles/profile_eval_detail.py [batch]"""
import ctypes
import json
import os
import sys

os.environ.setdefault("ARGUS_PROFILE_DETAIL", "1")
import torch

sys.path.insert(0, '.')
from argus_b200 import _lib
from argus_b200.models import NCameraCNN
from bench import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
torch.manual_seed(42)
model = NCameraCNN().to(dev).eval()
imgs, _ = synthetic_batch(B, 2, 256, 256, 0)
imgs = imgs.to(dev)
with torch.no_grad():
    for _ in range(3):
        model._forward_impl(imgs, False)
    lib = _lib.load()
    lib.argus_profile_enable(1)
    reps = 5
    for _ in range(reps):
        model._forward_impl(imgs, False)
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 18)
_lib.check(lib.argus_profile_report(buf, ctypes.c_int(1 << 18)))
fam = json.loads(buf.value.decode())
rows = []
for k, f in fam.items():
    n = f['launches'] / reps; ms = f['ms'] / reps
    tf = f['flops'] / reps / (ms / 1e3) / 1e12 if f['flops'] else 0
    gb = f['bytes'] / reps / (ms / 1e3) / 1e9 if f['bytes'] else 0
    rows.append((ms, k, n, tf, gb))
rows.sort(reverse=True)
print(f"batch {B}: total {sum(r[0] for r in rows):.3f} ms")
for ms, k, n, tf, gb in rows:
    print(f"{ms:8.4f} ms  x{n:4.0f}  {tf:7.1f} TF/s {gb:8.1f} GB/s  {k}")
