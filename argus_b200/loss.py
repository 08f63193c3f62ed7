"""geometric_loss_fn on the GPU (reference: /root/reference/argus/train.py:105-119)."""
from __future__ import annotations

import torch

from . import _lib


class _GeometricLoss(torch.autograd.Function):
    """One kernel computes the per-sample loss and its analytic gradient wrt pred; backward is a broadcast multiply."""

    @staticmethod
    def forward(ctx, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        p = pred.detach().to(torch.float32).reshape(-1, 6).contiguous()
        t = target.detach().to(torch.float32).reshape(-1, 7).contiguous()
        if p.shape[0] != t.shape[0]:
            raise ValueError(f"pred has {p.shape[0]} poses but target has {t.shape[0]}")
        n = p.shape[0]
        loss = torch.empty(n, dtype=torch.float32, device=p.device)
        grad = torch.empty((n, 6), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.call("argus_pose_loss", p, t, loss, None, grad, int(n), 1.0, _lib.stream_ptr())
        ctx.save_for_backward(grad)
        ctx.pred_shape = pred.shape
        ctx.pred_dtype = pred.dtype
        return loss.reshape(pred.shape[:-1])

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (grad,) = ctx.saved_tensors
        g = grad_out.reshape(-1, 1).to(torch.float32) * grad
        return g.reshape(ctx.pred_shape).to(ctx.pred_dtype), None


def geometric_loss_fn(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """The geometric loss function: squared norm of the SE(3) log-error between Exp(pred) and target.

    Args:
        pred: The predicted poses in se(3) of shape (..., 6), ordered [tau, phi].
        target: The target poses in SE(3) of shape (..., 7), ordered [t, qx, qy, qz, qw] (a pypose SE3 LieTensor
            is a Tensor subclass and is accepted as is).

    Returns:
        losses: The losses of shape (...).
    """
    if pred.shape[-1] != 6 or target.shape[-1] != 7:
        raise ValueError("pred must be (..., 6) se3 vectors and target (..., 7) SE3 poses")
    if not pred.is_cuda:
        raise _lib.ArgusError("geometric_loss_fn runs on sm_100a GPUs only (no CPU fallback)")
    target = target.as_subclass(torch.Tensor) if type(target) is not torch.Tensor else target
    return _GeometricLoss.apply(pred, target.to(pred.device))
