"""The reference's amp mode (argus/train.py:74,234,298-300,316-320: autocast + GradScaler) on the fused engine: loss
scaling, unscale before clip_grad_norm_, a non-finite step is skipped as a whole and backs the scale off."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(scaler, steps, poison_step=None):
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets, structured_images

    torch.manual_seed(0)
    model = NCameraCNN().to("cuda")
    eng = TrainEngine(model, lr=1e-3, max_grad_norm=1.0, distributed=False, scaler=scaler)
    losses, params = [], []
    for i in range(steps):
        x = structured_images(4, 6, 64, 64, 10 + i, "cuda")
        t = random_targets(4, 20 + i, "cuda")
        if i == poison_step:
            t = t.clone()
            t[0, 0] = float("inf")          # a non-finite loss -> non-finite gradients
        losses.append(float(eng.step(x, t)))
        params.append(model.flat_params.clone())
    return eng, losses, params


def test_loss_scaling_is_transparent(cuda_device):
    """Scaling by a power of two and unscaling in the optimizer changes no bit of the trajectory (bf16 and fp32 roundings
    are scale invariant away from overflow / underflow): the amp run equals the plain run."""
    from argus_b200.engine import GradScaler

    _, l0, p0 = _run(None, 4)
    eng, l1, p1 = _run(GradScaler(init_scale=2.0 ** 16), 4)
    assert l0 == l1
    assert all(torch.equal(a, b) for a, b in zip(p0, p1))
    assert eng.scaler.get_scale() == 2.0 ** 16 and eng.scaler.skipped_steps == 0 and eng.step_count == 4


def test_non_finite_step_is_skipped_and_backs_the_scale_off(cuda_device):
    from argus_b200.engine import GradScaler

    eng, losses, params = _run(GradScaler(init_scale=2.0 ** 16, growth_interval=2), 5, poison_step=2)
    assert not math.isfinite(losses[2])
    assert torch.equal(params[2], params[1])                       # nothing was updated by the poisoned step
    assert torch.isfinite(params[4]).all() and not torch.equal(params[3], params[2])
    assert eng.scaler.skipped_steps == 1 and eng.step_count == 4   # Adam's counter skipped it too
    # 2 clean steps -> growth (x2), poisoned step -> backoff (x0.5), 2 clean steps -> growth (x2)
    assert eng.scaler.get_scale() == 2.0 ** 17
    # a disabled scaler (amp=False, the reference default) is the plain path
    eng2, l2, _ = _run(GradScaler(enabled=False), 2)
    assert eng2.scaler.get_scale() == 1.0 and all(math.isfinite(v) for v in l2)


def test_train_config_amp_flag_builds_an_enabled_scaler(cuda_device, tmp_path):
    from argus_b200.engine import GradScaler

    s = GradScaler(enabled=True)
    assert s.is_enabled() and s.get_scale() == 65536.0
    x = torch.ones(3, device="cuda")
    assert torch.equal(s.scale(x), x * 65536.0)
    sd = s.state_dict()
    s2 = GradScaler()
    s2.load_state_dict(sd)
    assert s2.get_scale() == s.get_scale()
