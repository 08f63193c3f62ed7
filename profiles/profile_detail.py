import sys, json, ctypes, torch
sys.path.insert(0, '.')
from argus_b200 import _lib
from argus_b200.engine import TrainEngine
from argus_b200.models import NCameraCNN
from bench import synthetic_batch
dev = torch.device('cuda', 0)
torch.manual_seed(42)
model = NCameraCNN().to(dev)
eng = TrainEngine(model, distributed=False)
imgs, tgt = synthetic_batch(256, 2, 256, 256, 0)
imgs, tgt = imgs.to(dev), tgt.to(dev)
for _ in range(3): eng.step(imgs, tgt)
lib = _lib.load()
_lib.call('argus_model_set_wgrad_overlap', model._handle.ptr, 0)  # isolated per-kernel times
lib.argus_profile_enable(1)
for _ in range(2): eng.step(imgs, tgt)
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 18)
_lib.check(lib.argus_profile_report(buf, ctypes.c_int(1 << 18)))
fam = json.loads(buf.value.decode())
rows = []
for k, f in fam.items():
    n = f['launches'] / 2; ms = f['ms'] / 2
    tf = f['flops'] / 2 / (ms / 1e3) / 1e12 if f['flops'] else 0
    gb = f['bytes'] / 2 / (ms / 1e3) / 1e9 if f['bytes'] else 0
    rows.append((ms, k, n, tf, gb))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total {tot:.2f} ms")
for ms, k, n, tf, gb in rows:
    print(f"{ms:8.3f} ms  x{n:4.0f}  {tf:7.1f} TF/s {gb:8.1f} GB/s  {k}")
