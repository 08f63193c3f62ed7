"""Debug: uint8 input without augmentation through TrainEngine.step, synchronising after every C-ABI call."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200 import _lib  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
orig_call, orig_check = _lib.call, _lib.check


def call(name, *a):
    orig_call(name, *a)
    try:
        torch.cuda.synchronize()
    except Exception as e:   # noqa: BLE001
        print("FAILED after", name, [getattr(x, "shape", x) for x in a], flush=True)
        raise e
    print("ok", name, flush=True)


import os  # noqa: E402
if not os.environ.get("NOSYNC"):
    _lib.call = call
torch.manual_seed(42)
model = NCameraCNN().to(dev)
engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, augmentation=None)
batches = []
for k in range(3 if os.environ.get("MULTI") else 1):
    imgs, tgt = synthetic_batch(B, 2, 256, 256, seed=k)
    batches.append((imgs.to(dev), tgt.to(dev)))
for i in range(4):
    loss = engine.step(*batches[i % len(batches)])
    if not os.environ.get("NOSYNC") or os.environ.get("STEPSYNC"):
        torch.cuda.synchronize()
        print("step", i, float(loss), flush=True)
torch.cuda.synchronize()
print("done", float(loss))
