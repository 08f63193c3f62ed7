"""GPU: the fused clip_grad_norm_ + Adam step (argus_clip_adam_step) against torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam (reference argus/train.py:232,318-319), fp32 tolerance; and the engine's whole-step bookkeeping."""
import ctypes

import pytest
import torch

from argus_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,max_norm,gscale", [(1_000_003, 1.0, 1.0), (4096, 1e9, 0.5), (25_885_768, 1.0, 0.125)])
def test_clip_adam_matches_torch(cuda_device, n, max_norm, gscale):
    g = torch.Generator().manual_seed(n % 1000)
    p0 = torch.randn(n, generator=g).to(cuda_device)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    scratch = torch.zeros(1024, device=cuda_device)
    norm = torch.zeros(1, device=cuda_device)
    lib = _lib.load()
    for step in range(1, 5):
        grad = (torch.randn(n, generator=g) * (0.01 * step)).to(cuda_device)
        p_ref.grad = grad.clone() * gscale          # DDP hands the optimizer the averaged gradient
        ref_norm = torch.nn.utils.clip_grad_norm_([p_ref], max_norm)
        opt.step()
        _lib.check(lib.argus_clip_adam_step(_lib.ptr(p), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), ctypes.c_int64(n),
                                            _lib.ptr(scratch), ctypes.c_float(gscale), ctypes.c_float(max_norm),
                                            ctypes.c_float(1e-3), ctypes.c_float(0.9), ctypes.c_float(0.999),
                                            ctypes.c_float(1e-8), ctypes.c_int(step), _lib.ptr(norm), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert abs(norm.item() - ref_norm.item()) <= 2e-5 * ref_norm.item()
        assert torch.allclose(p, p_ref.detach(), rtol=2e-5, atol=2e-6), (p - p_ref.detach()).abs().max()
    state = opt.state[p_ref]
    assert torch.allclose(m, state["exp_avg"], rtol=1e-4, atol=1e-7)
    assert torch.allclose(v, state["exp_avg_sq"], rtol=1e-4, atol=1e-9)


def test_engine_gradients_equal_autograd_path(cuda_device):
    """TrainEngine.forward_backward (fused loss + staged backward into the flat arena) produces the same gradients as
    the drop-in autograd path `geometric_loss_fn(model(x), t).mean().backward()`, and is bitwise reproducible."""
    from argus_b200.engine import TrainEngine
    from argus_b200.loss import geometric_loss_fn
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets, structured_images

    torch.manual_seed(3)
    a = NCameraCNN().to(cuda_device)
    with torch.no_grad():
        for name, p in a.named_parameters():
            if name.endswith("bn3.weight"):
                p.fill_(0.1)
    b = NCameraCNN().to(cuda_device)
    b.load_state_dict(a.state_dict())
    x = structured_images(8, 6, 128, 128, 9, cuda_device)
    t = random_targets(8, 10, cuda_device)
    eng = TrainEngine(a, lr=1e-3, max_grad_norm=1.0, distributed=False)
    loss_a = eng.forward_backward(x, t)
    ga = a.flat_grads.clone()
    loss_a2 = eng.forward_backward(x, t)            # same path again: every reduction is ordered, so the result is
    ga2 = a.flat_grads.clone()                      # bitwise identical (BN running statistics do not enter train mode)
    assert torch.equal(ga, ga2) and loss_a.item() == loss_a2.item()
    noise = 0.0
    b.train()
    loss_b = geometric_loss_fn(b(x), t).mean()
    loss_b.backward()
    gb = torch.zeros_like(ga)
    for p, (_n, off, numel, _shape) in zip(b.parameters(), b._param_infos):
        gb[off:off + numel] = p.grad.reshape(-1)
    assert abs(loss_a.item() - loss_b.item()) < 1e-3 * abs(loss_b.item())
    rel = ((ga - gb).norm() / gb.norm()).item()
    print(f"engine vs autograd path: {rel:.3e}; engine run-to-run: {noise:.3e}")
    assert rel < 1e-6, (rel, noise)
    eng.optimizer_step()
    assert eng.step_count == 1 and int(a.resnet.bn1.num_batches_tracked) == 2  # two training forwards so far
    assert torch.isfinite(a.flat_params).all()


def test_head_and_fc_against_torch(cuda_device):
    """Avg-pool -> fc -> GELU -> MLP head in isolation: feed identical layer4 activations through a torch fp32 head."""
    from argus_b200.models import NCameraCNN
    from oracle.ref_model import make_reference_model

    ref = make_reference_model(5).to(cuda_device).eval()
    ours = NCameraCNN().to(cuda_device).eval()
    ours.load_state_dict(ref.state_dict())
    x = torch.rand(3, 6, 64, 64, device=cuda_device)
    with torch.no_grad():
        ours(x)
        feat = ours.probe_activation(15).float()            # (N*HW, 2048) layer4 output of OUR network
        pooled = feat.reshape(6, -1, 2048).mean(1)
        f = ref.resnet.fc(pooled.bfloat16().float()).reshape(3, 2048)
        want = ref.output_mlp(torch.nn.functional.gelu(f))
        got = ours(x)
    assert torch.allclose(got, want, rtol=2e-2, atol=2e-3), (got, want)
