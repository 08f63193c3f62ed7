"""Experiment: repeat a short training trajectory from the same seed at the bench batch size and report the first
step whose loss differs from the first trial's. Modes: aug 0|1, prefetch 0|1 (look-ahead staging on the side stream).
Usage: python profiles/experiments/trajectory_repeat.py <aug> <prefetch> [trials] [steps] [batch]"""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200.data import Augmentation, AugmentationConfig  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

aug_on, prefetch = int(sys.argv[1]), int(sys.argv[2])
trials = int(sys.argv[3]) if len(sys.argv) > 3 else 8
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
B = int(sys.argv[5]) if len(sys.argv) > 5 else 256
GAP_AFTER = int(os.environ.get("GAP_AFTER", "-1"))
GAP_S = float(os.environ.get("GAP_S", "1.0"))
dev = torch.device("cuda", 0)
batches = []
for k in range(3):
    imgs, tgt = synthetic_batch(B, 2, 256, 256, seed=k)
    batches.append((imgs.to(dev), tgt.to(dev)))
torch.cuda.synchronize()
# POISON=<nan|zero|rand>: fill (most of) the free device memory with a pattern and hand it back to the driver, so that
# every cudaMalloc of the library returns pages with known garbage: a read of uninitialised memory shows up as a NaN
# or as a result that depends on the pattern
poison = os.environ.get("POISON")
if poison:
    free, _total = torch.cuda.mem_get_info(dev)
    n = int(free * 0.9) // 4
    junk = torch.empty(n, dtype=torch.float32, device=dev)
    if poison == "nan":
        junk.fill_(float("nan"))
    elif poison == "zero":
        junk.zero_()
    else:
        junk.uniform_(-3.0, 3.0)
    torch.cuda.synchronize()
    del junk
    torch.cuda.empty_cache()
ref = None
bad = 0
for t in range(trials):
    torch.manual_seed(42)
    model = NCameraCNN().to(dev)
    aug = Augmentation(AugmentationConfig(), train=True, seed=1, gpu_spaghetti=True) if aug_on else None
    engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, distributed=False, augmentation=aug)
    losses = []
    for i in range(steps):
        losses.append(engine.step(*batches[i % 3]).clone())
        if prefetch and i >= GAP_AFTER:
            engine.prefetch(batches[(i + 1) % 3][0])
        if i == GAP_AFTER:
            torch.cuda.synchronize()
            time.sleep(GAP_S)      # idle GPU: clocks drop, the next kernels run with different timing (bench.py waits here)
    torch.cuda.synchronize()
    cur = [float(x) for x in losses]
    if ref is None:
        ref = cur
    elif cur != ref:
        bad += 1
        first = next(i for i, (a, b) in enumerate(zip(cur, ref)) if a != b)
        print(f"trial {t}: first differing step {first}: {cur[first]!r} vs {ref[first]!r}", flush=True)
    del engine, model
print(f"aug={aug_on} prefetch={prefetch}: {bad} of {trials - 1} trials differ from the first")
print("LOSSES", " ".join(repr(x) for x in ref))
