"""Experiment: per-tensor gradient checksums of one train-mode forward/backward (B=8, 128x128), repeated, to localise
an ordering bug. Usage: python profiles/experiments/pdl_race.py <overlap 0|1> [reps]   (env: ARGUS_PDL, ARGUS_B200_LIB)"""
import hashlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from gpu_util import random_targets, structured_images  # noqa: E402
from argus_b200 import _lib  # noqa: E402
from argus_b200.loss import geometric_loss_fn  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402

overlap = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = NCameraCNN().to(dev)
with torch.no_grad():
    for m in model.modules():
        if hasattr(m, "bn3"):
            m.bn3.weight.fill_(0.1)
model.train()
x = structured_images(8, 6, 128, 128, 3, dev)
target = random_targets(8, 4, dev)
model(x)   # create the native handle
_lib.call("argus_model_set_wgrad_overlap", model._handle.ptr, overlap)
first = None
for r in range(reps):
    model.zero_grad(set_to_none=True)
    y = model(x)
    loss = geometric_loss_fn(y, target).mean()
    loss.backward()
    torch.cuda.synchronize()
    sums = {n: hashlib.md5(p.grad.cpu().numpy().tobytes()).hexdigest()[:8] for n, p in model.named_parameters()}
    if first is None:
        first = sums
        print("rep 0 loss %.6f  all %s" % (loss.item(), hashlib.md5("".join(sums.values()).encode()).hexdigest()[:12]))
        for n in list(sums)[:3] + list(sums)[-3:]:
            print("   ", n, sums[n])
    else:
        diff = [n for n in sums if sums[n] != first[n]]
        print("rep %d loss %.6f  differing tensors: %d %s" % (r, loss.item(), len(diff), diff[:6]))
