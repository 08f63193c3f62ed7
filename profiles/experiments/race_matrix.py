"""Experiment: which ingredient makes the training trajectory irreproducible when kernels trigger their dependents
early (-DARGUS_PDL_TRIGGER build + ARGUS_PDL=1)?  Many short trials in ONE process (no start-up cost per trial):
every trial rebuilds model + engine from the same seed, runs `steps` training steps with look-ahead staging, and
records per step (without any host synchronisation inside the trajectory)
    in   : checksum of the pooled stem output  -> depends only on the staged (augmented) input and the stem weights
    out  : the mean loss                       -> the forward pass
    grad : checksum of the flat gradient arena -> the backward pass
    par  : checksum of the parameters after the optimizer
The first trial of the first configuration is the reference; a trial "diverges" at the first (step, quantity) whose
checksum differs. Configurations vary the caller's stream (legacy default / plain non-default / high priority) and the
look-ahead staging; the library build and ARGUS_PDL* come from the environment (one process per build).
Usage: python profiles/experiments/race_matrix.py <out.json> [trials] [steps] [batch]"""
import ctypes
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200 import _lib  # noqa: E402
from argus_b200.data import Augmentation, AugmentationConfig  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

out_path = sys.argv[1]
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)

batches = []
for k in range(3):
    imgs, tgt = synthetic_batch(B, 2, 256, 256, seed=k)
    batches.append((imgs.to(dev), tgt.to(dev)))
torch.cuda.synchronize()
QUANT = ["in", "out", "grad", "par", "par_before"]
BLOCKS = os.environ.get("RACE_BLOCKS") == "1"     # also checksum the 16 bottleneck outputs, pooled features, fc output
if BLOCKS:
    QUANT = QUANT + [f"block{i}" for i in range(16)] + ["avgpool", "fc"]
LIGHT = os.environ.get("RACE_LIGHT") == "1"


def csum(t):
    # order-independent of launch configuration: a double-precision sum of |x| plus a sum of x * ramp (position-sensitive)
    f = t.detach().reshape(-1).double()
    return f.abs().sum() + (f * torch.arange(f.numel(), device=f.device, dtype=torch.float64).remainder(977.0)).sum()


def trial(stream_mode, prefetch):
    if stream_mode == "default":
        stream = torch.cuda.default_stream(dev)
    elif stream_mode == "plain":
        stream = torch.cuda.Stream(device=dev)
    else:
        stream = torch.cuda.Stream(device=dev, priority=-1)
    sig = torch.zeros(steps, len(QUANT), dtype=torch.float64, device=dev)
    with torch.cuda.stream(stream):
        torch.manual_seed(42)
        model = NCameraCNN().to(dev)
        aug = Augmentation(AugmentationConfig(), train=True, seed=7).to(dev)
        engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, augmentation=aug, distributed=False)
        pooled = torch.empty(B * 2 * 64 * 64 * 64, dtype=torch.bfloat16, device=dev)
        big = torch.empty(B * 2 * 64 * 64 * 256, dtype=torch.bfloat16, device=dev) if BLOCKS else None
        model._ensure_bound()
        for i in range(steps):
            sig[i, 4] = csum(model.flat_params) + csum(model._flat_buffers)   # before the step: initialisation / H2D
            if LIGHT:
                # nothing but the training loop's own calls between the steps (what tests / bench.py do)
                loss = engine.step(*batches[i % 3])
                if prefetch:
                    engine.prefetch(batches[(i + 1) % 3][0])
                sig[i, 1] = loss.double()
                if i == steps - 1:
                    sig[i, 3] = csum(model.flat_params)
                continue
            engine.forward_backward(*batches[i % 3])
            if prefetch:
                engine.prefetch(batches[(i + 1) % 3][0])   # look-ahead staging enqueued before anything is inspected
            with torch.cuda.device(dev):
                _lib.check(_lib.load().argus_model_copy_activation(
                    model._handle.ptr, ctypes.c_int(-1), _lib.ptr(pooled), ctypes.c_int64(pooled.numel()), None, None,
                    _lib.stream_ptr()))
            sig[i, 0] = csum(pooled)
            if BLOCKS:
                for j in range(18):
                    rows, C = ctypes.c_int64(), ctypes.c_int()
                    with torch.cuda.device(dev):
                        _lib.check(_lib.load().argus_model_copy_activation(
                            model._handle.ptr, ctypes.c_int(j), _lib.ptr(big), ctypes.c_int64(big.numel()),
                            ctypes.byref(rows), ctypes.byref(C), _lib.stream_ptr()))
                    sig[i, 5 + j] = csum(big[: rows.value * C.value])
            sig[i, 1] = engine._loss_mean[0].double()     # the forward pass (network output -> loss)
            sig[i, 2] = csum(model.flat_grads)
            engine.optimizer_step()
            sig[i, 3] = csum(model.flat_params)
    torch.cuda.synchronize()
    del engine, model
    return sig.cpu()


configs = [("default", 1), ("prio", 1), ("prio", 0), ("plain", 1)]
if os.environ.get("RACE_CONFIGS"):   # e.g. "default:1,prio:0"
    configs = [(c.split(":")[0], int(c.split(":")[1])) for c in os.environ["RACE_CONFIGS"].split(",")]
ref = None
report = {"env": {k: v for k, v in os.environ.items() if k.startswith("ARGUS_")}, "trials": trials, "steps": steps,
          "batch": B, "configs": []}
all_sigs = []
for (mode, pf) in configs:
    div = []
    for t in range(trials):
        s = trial(mode, pf)
        all_sigs.append({"stream": mode, "prefetch": pf, "trial": t, "sig": [[repr(float(v)) for v in row] for row in s]})
        if ref is None:
            ref = s
            continue
        bad = (s != ref).nonzero()
        if bad.numel():
            st, q = int(bad[0, 0]), int(bad[0, 1])
            div.append({"trial": t, "step": st, "first": QUANT[q], "all_in_step": [QUANT[int(b[1])] for b in bad if int(b[0]) == st]})
            print("   divergent:", json.dumps(div[-1]), flush=True)
    rec = {"stream": mode, "prefetch": pf, "diverged": len(div), "of": trials - (1 if (mode, pf) == configs[0] else 0),
           "detail": div[:12]}
    report["configs"].append(rec)
    print(json.dumps(rec), flush=True)
report["signatures"] = all_sigs
# distinct trajectories over ALL trials of the process (1 = every trial reproduced the same bits)
keys = {}
for a in all_sigs:
    keys.setdefault(json.dumps(a["sig"]), []).append((a["stream"], a["prefetch"], a["trial"]))
report["distinct_trajectories"] = sorted((len(v) for v in keys.values()), reverse=True)
print("distinct trajectories (cluster sizes):", report["distinct_trajectories"], flush=True)
json.dump(report, open(out_path, "w"), indent=1)
