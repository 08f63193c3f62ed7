"""ORACLE (test infrastructure only): CPU/PyTorch restatement of the reference model.

Follows /root/reference/argus/models.py line by line:
  :43      self.resnet = torchvision.models.resnet50(weights="DEFAULT")   -> weights=None here (no network; the
           north-star asks for random-init weights), everything else unchanged
  :55-56   avgpool -> AdaptiveAvgPool2d((1,1)); fc -> Linear(2048, resnet_output_dim)
  :58-64   output_mlp = Linear(n_cams*dim,128) GELU Linear(128,128) GELU Linear(128,6)
  :76-90   forward: assert 4-D; reshape(-1,3,H,W); resnet; reshape(B, n_cams*dim); GELU; output_mlp

torchvision (third-party, 0.26.0 installed in this image; the reference pins only `torchvision>=0.15.2`,
pyproject.toml) supplies the ResNet-50 definition exactly as it does for the reference.

Pinning: `oracle/make_golden.py` (run in the build container, where /root/reference exists) asserts that this class
and the real `argus.models.NCameraCNN` (imported from /root/reference with the weight download shimmed out) have
identical state_dict keys/shapes and bit-identical outputs and gradients under the same seed, then writes
tests/golden/*.json. tests/test_oracle_model.py re-checks those vectors anywhere.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torchvision.models as tvm


class RefNCameraCNN(nn.Module):
    def __init__(self, n_cams: int = 2, resnet_output_dim: int = 1024) -> None:
        super().__init__()
        self.resnet = tvm.resnet50(weights=None)
        self.num_channels = 3 * n_cams
        self.resnet_output_dim = resnet_output_dim
        self.n_cams = n_cams
        self.resnet.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.resnet.fc = nn.Linear(self.resnet.fc.in_features, self.resnet_output_dim)
        self.output_mlp = nn.Sequential(
            nn.Linear(self.n_cams * self.resnet_output_dim, 128),
            nn.GELU(),
            nn.Linear(128, 128),
            nn.GELU(),
            nn.Linear(128, 6),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert len(x.shape) == 4, "The input images must be of shape (B, C, H, W)! If B=1, add a dummy dimension."
        B = x.shape[0]
        x = x.reshape(-1, 3, *(x.shape[-2:]))
        x = self.resnet(x)
        x = x.reshape(B, self.n_cams * self.resnet_output_dim)
        x = nn.GELU()(x)
        return self.output_mlp(x)


def make_reference_model(seed: int = 42, n_cams: int = 2, resnet_output_dim: int = 1024) -> RefNCameraCNN:
    """Reference-initialised model under the reference's default seed (argus/train.py:70,127-128)."""
    torch.manual_seed(seed)
    return RefNCameraCNN(n_cams, resnet_output_dim)


def torch_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Differentiable torch (float64) version of oracle.se3_loss.geometric_loss for end-to-end gradient oracles.
    Uses the closed forms away from the series region (inputs in tests are generic)."""
    pred = pred.double()
    target = target.double()
    tau, phi = pred[..., :3], pred[..., 3:]
    th = phi.norm(dim=-1, keepdim=True)

    def hat(v):
        z = torch.zeros_like(v[..., 0])
        return torch.stack([torch.stack([z, -v[..., 2], v[..., 1]], -1), torch.stack([v[..., 2], z, -v[..., 0]], -1),
                            torch.stack([-v[..., 1], v[..., 0], z], -1)], -2)

    def qmul(a, b):
        av, aw, bv, bw = a[..., :3], a[..., 3:], b[..., :3], b[..., 3:]
        return torch.cat([aw * bv + bw * av + torch.cross(av, bv, dim=-1), aw * bw - (av * bv).sum(-1, keepdim=True)], -1)

    def qrot(q, x):
        v, w = q[..., :3], q[..., 3:]
        t = 2 * torch.cross(v, x, dim=-1)
        return x + w * t + torch.cross(v, t, dim=-1)

    K = hat(phi)
    th2 = th[..., None]
    Jl = torch.eye(3, dtype=torch.float64, device=pred.device) + (1 - torch.cos(th2)) / th2**2 * K + (th2 - torch.sin(th2)) / th2**3 * (K @ K)
    qp = torch.cat([phi * torch.sin(th / 2) / th, torch.cos(th / 2)], -1)
    tp = (Jl @ tau[..., None])[..., 0]
    t, q = target[..., :3], target[..., 3:]
    qc = torch.cat([-q[..., :3], q[..., 3:]], -1)
    tinv = -qrot(qc, t)
    qe = qmul(qp, qc)
    te = tp + qrot(qp, tinv)
    v, w = qe[..., :3], qe[..., 3:]
    n = v.norm(dim=-1, keepdim=True)
    phie = v * (2 * torch.atan(n / w) / n)
    the = phie.norm(dim=-1, keepdim=True)[..., None]
    Ke = hat(phie)
    Jinv = torch.eye(3, dtype=torch.float64, device=pred.device) - 0.5 * Ke + (1 / the**2 - (1 + torch.cos(the)) / (2 * the * torch.sin(the))) * (Ke @ Ke)
    taue = (Jinv @ te[..., None])[..., 0]
    return (taue**2).sum(-1) + (phie**2).sum(-1)
