"""Profiling target: N fused training steps at the bench configuration (B=256 pairs, 256x256, augmentation on).
Usage under ncu: see profiles/README.md. Prints the number of library launches per step."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from argus_b200 import _lib  # noqa: E402
from argus_b200.data import Augmentation, AugmentationConfig  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = NCameraCNN().to(dev)
eng = TrainEngine(model, distributed=False, augmentation=Augmentation(AugmentationConfig(), train=True, seed=1, gpu_spaghetti=True))
imgs, tgt = synthetic_batch(batch, 2, 256, 256, 0)
imgs, tgt = imgs.to(dev), tgt.to(dev)
lib = _lib.load()
lib.argus_launch_count.restype = ctypes.c_int64
n0 = lib.argus_launch_count()
for _ in range(steps):
    loss = eng.step(imgs, tgt)
torch.cuda.synchronize()
print("launch scopes per step:", (lib.argus_launch_count() - n0) / steps, "loss", float(loss))
