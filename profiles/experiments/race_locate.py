"""Experiment: locate the first tensor that differs between a default-stream run and a high-priority-stream run with
look-ahead staging. Per step: md5 of the pooled stem output (depends only on the staged input and the stem weights), of
the flat gradient arena after backward, and of the parameters after the optimizer.
Usage: python profiles/experiments/race_locate.py <default|prio> <out.json> [steps]"""
import hashlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200.data import Augmentation, AugmentationConfig  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

mode, out = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if mode == "prio":
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
torch.manual_seed(42)
model = NCameraCNN().to(dev)
engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, augmentation=Augmentation(AugmentationConfig(), train=True).to(dev))
batches = []
for k in range(3):
    imgs, tgt = synthetic_batch(64, 2, 256, 256, seed=k)
    batches.append((imgs.to(dev), tgt.to(dev)))


def h(t):
    return hashlib.md5(t.detach().contiguous().view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:10]


rec = []
for i in range(steps):
    loss = engine.forward_backward(*batches[i % 3])
    # enqueue the look-ahead staging BEFORE reading anything back, exactly like the training loop
    engine.prefetch(batches[(i + 1) % 3][0])
    g = model.flat_grads.clone()
    pooled = model.probe_activation(-1).clone()
    engine.optimizer_step()
    p = model.flat_params.clone()
    per = {n: h(g[off:off + numel]) for (n, off, numel, _shape) in model._param_infos}
    rec.append({"step": i, "loss": float(loss), "pooled": h(pooled), "grads": h(g), "params": h(p), "per": per})
torch.cuda.synchronize()
json.dump(rec, open(out, "w"), indent=0)
print(mode, [r["loss"] for r in rec][-3:])
