"""North star: "bf16 mode <= 2e-2 relative, with a matching 1k-step loss curve".

Trains the SAME network from the SAME initial weights on the SAME batch sequence four ways and compares the curves:
  ref32   the reference model (oracle/ref_model.py == /root/reference/argus/models.py) in PyTorch fp32 (TF32 off),
          torch.optim.Adam(1e-4) + clip_grad_norm_(1.0): the step body of /root/reference/argus/train.py:298-320
  ref16   the same under torch.autocast(bfloat16) -- what bf16 costs the REFERENCE's own stack
  ours16  argus_b200 TrainEngine, bf16 tensor-core path (the product)
  ours32  argus_b200 TrainEngine in the fp32 parity mode
The task is learnable: 64 synthetic pairs whose low-frequency image content encodes the target pose, visited in a fixed
order in batches of 8; no augmentation (its RNG differs by design). Usage:
    python profiles/loss_curve.py [steps] [out.json]
"""
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def make_task(n=64, size=128, seed=0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    target = torch.cat([0.3 * torch.randn(n, 3, generator=g), q / q.norm(dim=-1, keepdim=True)], -1)
    # the pose drives a low-frequency pattern in both views, plus per-sample texture noise
    basis = torch.randn(7, 6, 8, 8, generator=g)
    coarse = torch.einsum("nk,kchw->nchw", target, basis)
    img = F.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
    img = (0.5 + 0.25 * img + 0.05 * torch.randn(n, 6, size, size, generator=g)).clamp(0, 1)
    return img.to(device), target.to(device)


def run_reference(images, targets, steps, batch, autocast, model_seed=42, channels_last=False, tf32=False):
    """The reference's step body (argus/train.py:298-320) on the reference model. channels_last / tf32 select other
    cuDNN kernels for the SAME fp32 math (different summation order / the TF32 convolutions PyTorch uses by default on
    Ampere and later): the spread between these runs is the reference's own run-to-run numerical spread."""
    from oracle.ref_model import make_reference_model, torch_loss

    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    model = make_reference_model(model_seed).to(images.device).train()
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    losses = []
    n = images.shape[0]
    for s in range(steps):
        i0 = (s * batch) % n
        x, t = images[i0:i0 + batch], targets[i0:i0 + batch]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = model(x)
        loss = torch_loss(y.float(), t).mean().float()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        losses.append(loss.detach())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.stack(losses).cpu().tolist()


def run_ours(images, targets, steps, batch, precision, model_seed=42):
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from oracle.ref_model import make_reference_model

    ref = make_reference_model(model_seed)
    model = NCameraCNN().to(images.device).set_precision(precision)
    model.load_state_dict(ref.state_dict())
    eng = TrainEngine(model, lr=1e-4, max_grad_norm=1.0, distributed=False)
    losses = []
    n = images.shape[0]
    for s in range(steps):
        i0 = (s * batch) % n
        losses.append(eng.step(images[i0:i0 + batch], targets[i0:i0 + batch]).clone())
    return torch.stack(losses).cpu().tolist()


def window_means(curve, w):
    return [sum(curve[i:i + w]) / w for i in range(0, len(curve) - w + 1, w)]


def compare(steps=1000, batch=8, size=128, window=100):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    images, targets = make_task(size=size)
    curves = {
        "ref32": run_reference(images, targets, steps, batch, autocast=False),
        "ref16": run_reference(images, targets, steps, batch, autocast=True),
        "ours16": run_ours(images, targets, steps, batch, "bf16"),
        "ours32": run_ours(images, targets, steps, batch, "fp32"),
    }
    means = {k: window_means(v, window) for k, v in curves.items()}
    rel = {k: [abs(a - b) / b for a, b in zip(means[k], means["ref32"])] for k in ("ref16", "ours16", "ours32")}
    return {"steps": steps, "batch": batch, "size": size, "window": window, "window_means": means,
            "rel_to_ref32": rel, "first_loss": {k: v[0] for k, v in curves.items()},
            "final_window": {k: v[-1] for k, v in means.items()}, "curves": curves}


def acceptance(steps=1000, batch=8, size=128, window=100, seeds=(0, 1, 2), slack=0.0):
    """Loss-curve acceptance (north star: "a matching 1k-step loss curve"): for every seed (task + initial weights), the
    reference is run three ways -- fp32, fp32 channels_last, TF32 (PyTorch's default convolution arithmetic) -- and every
    `window`-step mean of the bf16 product run must lie inside [min, max] of the three reference runs, widened by
    `slack` (relative). Returns the per-seed window means and, per window, where ours sits relative to the band."""
    out = {"steps": steps, "batch": batch, "size": size, "window": window, "slack": slack, "seeds": {}}
    worst = 0.0
    for seed in seeds:
        images, targets = make_task(size=size, seed=seed)
        refs = {"fp32": run_reference(images, targets, steps, batch, False, model_seed=42 + seed),
                "fp32_channels_last": run_reference(images, targets, steps, batch, False, model_seed=42 + seed,
                                                    channels_last=True),
                "tf32": run_reference(images, targets, steps, batch, False, model_seed=42 + seed, tf32=True)}
        ours = run_ours(images, targets, steps, batch, "bf16", model_seed=42 + seed)
        means = {k: window_means(v, window) for k, v in refs.items()}
        means["ours_bf16"] = window_means(ours, window)
        lo = [min(means[k][i] for k in refs) for i in range(len(means["ours_bf16"]))]
        hi = [max(means[k][i] for k in refs) for i in range(len(means["ours_bf16"]))]
        # signed distance outside the band, relative to the band edge (0 = inside)
        outside = [max(l / o - 1.0, o / h - 1.0, 0.0) for o, l, h in zip(means["ours_bf16"], lo, hi)]
        worst = max(worst, max(outside))
        out["seeds"][str(seed)] = {"window_means": means, "band_lo": lo, "band_hi": hi, "outside": outside,
                                   "ref_spread": [h / l - 1.0 for l, h in zip(lo, hi)]}
    out["worst_outside"] = worst
    out["accepted"] = worst <= slack
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "acceptance":
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
        out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/loss_curve_acceptance.json"
        res = acceptance(steps)
        for seed, r in res["seeds"].items():
            for k, v in r["window_means"].items():
                print(seed, f"{k:20s}", " ".join(f"{m:8.4f}" for m in v))
            print(seed, f"{'outside band':20s}", " ".join(f"{m:8.4f}" for m in r["outside"]))
            print(seed, f"{'reference spread':20s}", " ".join(f"{m:8.4f}" for m in r["ref_spread"]))
        print("worst outside", res["worst_outside"])
        Path(out).write_text(json.dumps(res))
        sys.exit(0)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/loss_curve.json"
    res = compare(steps)
    for k in ("ref32", "ref16", "ours16", "ours32"):
        print(k, " ".join(f"{m:8.4f}" for m in res["window_means"][k]))
    for k, v in res["rel_to_ref32"].items():
        print("rel", k, " ".join(f"{m:8.4f}" for m in v))
    Path(out).write_text(json.dumps(res))
