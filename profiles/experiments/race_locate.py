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


infos = list(model._param_infos)
sig = torch.zeros(steps, len(infos), dtype=torch.float64, device=dev)   # per-parameter gradient checksums, no host sync
pooled_sig = torch.zeros(steps, dtype=torch.float64, device=dev)
losses = []
for i in range(steps):
    loss = engine.forward_backward(*batches[i % 3])
    engine.prefetch(batches[(i + 1) % 3][0])   # look-ahead staging enqueued before anything is read back
    g = model.flat_grads
    sums = torch.stack([g[off:off + numel].double().abs().sum() for (_n, off, numel, _s) in infos])
    sig[i] = sums
    engine.optimizer_step()
    losses.append(loss.clone())
torch.cuda.synchronize()
rec = {"names": [n for (n, _o, _k, _s) in infos], "sig": sig.cpu().tolist(), "loss": [float(l) for l in losses]}
json.dump(rec, open(out, "w"))
print(mode, rec["loss"][-3:])
