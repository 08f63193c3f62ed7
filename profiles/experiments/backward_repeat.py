"""Experiment: the same forward + backward repeated on the same weights and batch must give the same bits.
Prints, for every repetition that differs from the first, the parameters whose gradient differs (network order).
Usage: python profiles/experiments/backward_repeat.py [reps] [batch] [noise 0|1|2]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402
from bench import synthetic_batch  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = NCameraCNN().to(dev)
engine = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, distributed=False)
imgs, tgt = synthetic_batch(B, 2, 256, 256, seed=0)
imgs, tgt = imgs.to(dev), tgt.to(dev)
infos = list(model._param_infos)
ref = None
ref_loss = None
n_bad = 0
# optional interference: a copy stream hammering HBM while the step runs (changes every latency inside the kernels)
NOISE = int(sys.argv[3]) if len(sys.argv) > 3 else 0
noise_stream = torch.cuda.Stream(device=dev)
na = torch.empty(1 << 28, dtype=torch.float32, device=dev)
nb = torch.empty_like(na)
for it in range(reps):
    if NOISE:
        noise_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(noise_stream):
            for _ in range(2 + it % 5):
                if NOISE == 1:
                    nb.copy_(na)            # copy engine / HBM only
                else:
                    torch.sin(na, out=nb)   # SM-resident elementwise CTAs: delays the CTA scheduling of the step's kernels
    loss = engine.forward_backward(imgs, tgt).clone()
    g = model.flat_grads.clone()
    torch.cuda.synchronize()
    if ref is None:
        ref, ref_loss = g, loss
        continue
    if not torch.equal(g, ref) or not torch.equal(loss, ref_loss):
        n_bad += 1
        bad = [(n, float((g[o:o + k] - ref[o:o + k]).abs().max() / (ref[o:o + k].abs().max() + 1e-30)))
               for (n, o, k, _s) in infos if not torch.equal(g[o:o + k], ref[o:o + k])]
        print(f"rep {it}: loss {'same' if torch.equal(loss, ref_loss) else 'DIFFERS'}; {len(bad)} of {len(infos)} gradients differ; "
              f"last (deepest) five: {bad[-5:]}", flush=True)
print(f"{n_bad} of {reps - 1} repetitions differ from the first")
