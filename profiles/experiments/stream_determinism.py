"""Experiment: the training trajectory must not depend on which CUDA stream the caller uses.
Usage: python profiles/experiments/stream_determinism.py <default|plain|prio> [prefetch 0|1] [aug 0|1]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "default"
prefetch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
aug_on = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if mode == "plain":
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
elif mode == "prio":
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
torch.manual_seed(42)
model = NCameraCNN().to(dev)
aug = None
if aug_on:
    from argus_b200.data import Augmentation, AugmentationConfig
    aug = Augmentation(AugmentationConfig(), train=True).to(dev)
engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, augmentation=aug)
B = 64
batches = []
for k in range(3):
    imgs, tgt = synthetic_batch(B, 2, 256, 256, seed=k)
    batches.append((imgs.to(dev), tgt.to(dev)))
losses = []
for i in range(8):
    losses.append(engine.step(*batches[i % 3]).clone())
    if prefetch:
        engine.prefetch(batches[(i + 1) % 3][0])
torch.cuda.synchronize()
print(mode, prefetch, aug_on, " ".join("%.9f" % float(l) for l in losses))
