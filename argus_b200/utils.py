"""Pose utilities of the hot path (reference: /root/reference/argus/utils.py:110-145,153-171,179-189)."""
from __future__ import annotations

from typing import Callable

import torch

from . import _lib


def xyzwxyz_to_xyzxyzw_SE3(xyzwxyz: torch.Tensor) -> torch.Tensor:
    """Converts 7d poses with quats from (w, x, y, z) to (x, y, z, w) order (reference: argus/utils.py:110-126)."""
    return torch.cat((xyzwxyz[..., :3], xyzwxyz[..., -3:], xyzwxyz[..., -4:-3]), dim=-1)


def xyzxyzw_to_xyzwxyz_SE3(xyzxyzw: torch.Tensor) -> torch.Tensor:
    """Converts 7d poses with quats from (x, y, z, w) to (w, x, y, z) order (reference: argus/utils.py:129-145)."""
    return torch.cat((xyzxyzw[..., :3], xyzxyzw[..., -1:], xyzxyzw[..., -4:-1]), dim=-1)


def se3_exp(pred: torch.Tensor, wxyz: bool = False) -> torch.Tensor:
    """`pp.se3(pred).Exp()` on the GPU: (..., 6) se3 [tau, phi] -> (..., 7) [t, qx, qy, qz, qw].

    With wxyz=True the quaternion is emitted scalar-first, i.e. xyzxyzw_to_xyzwxyz_SE3 is fused in."""
    if not pred.is_cuda:
        raise _lib.ArgusError("se3_exp runs on sm_100a GPUs only (no CPU fallback)")
    p = pred.detach().to(torch.float32).reshape(-1, 6).contiguous()
    out = torch.empty((p.shape[0], 7), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.call("argus_pose_exp", p, out, int(p.shape[0]), int(wxyz), _lib.stream_ptr())
    return out.reshape(*pred.shape[:-1], 7)


def get_pose(images: torch.Tensor, model: torch.nn.Module) -> torch.Tensor:
    """Get the pose of the cube from the images (reference: argus/utils.py:179-189).

    Args:
        images: The images of shape (B, 3 * n_cams, W, H), concatenated along the channel dimension.
        model: The model to use.

    Returns:
        pose: The predicted pose as a 7d pose (x, y, z, qx, qy, qz, qw). The reference returns a pypose SE3
            LieTensor; pypose is not a dependency here, so a plain (B, 7) tensor with the same data is returned.
    """
    return se3_exp(model(images))


def time_torch_fn(fn: Callable[[], torch.Tensor]) -> tuple[torch.Tensor, float]:
    """Time a torch function with CUDA events (reference: argus/utils.py:153-171). Returns (result, seconds)."""
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    start.record()
    result = fn()
    end.record()
    torch.cuda.synchronize()
    return result, start.elapsed_time(end) / 1000


def draw_spaghetti(img, n_arcs: int = 10, width_range=(1.0, 5.0)):
    """Draws random black arcs on a PIL image (reference: argus/utils.py:252-275; CPU/PIL, uses numpy's global RNG
    exactly like the reference so that `np.random.seed` reproduces it)."""
    import numpy as np
    from PIL import ImageDraw

    for _ in range(n_arcs):
        x0, y0 = np.random.randint(0, img.width), np.random.randint(0, img.height)
        x1, y1 = np.random.randint(x0, img.width), np.random.randint(y0, img.height)
        start_angle, end_angle = np.random.randint(0, 360), np.random.randint(0, 360)
        width = np.random.uniform(*width_range)
        d = ImageDraw.Draw(img)
        d.arc((x0, y0, x1, y1), start_angle, end_angle, fill=(0, 0, 0), width=int(width))
    return img
