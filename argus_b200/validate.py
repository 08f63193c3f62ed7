"""Validation entry point (reference: /root/reference/argus/validate.py:48-181). The compute part — load a `.pth`,
eval-mode forward at batch 1, geometric loss, pose exponential — runs on the B200 path; the reference's matplotlib
figures are host-side visualisation and out of scope: `validate()` returns the numbers instead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .data import AugmentationConfig
from .dataset import CameraCubePoseDataset, CameraCubePoseDatasetConfig
from .loss import geometric_loss_fn
from .models import NCameraCNN, NCameraCNNConfig
from .utils import se3_exp


@dataclass(frozen=True)
class ValConfig:
    """Configuration for validation (reference fields: validate.py:48-82)."""

    model_path: str
    dataset_config: CameraCubePoseDatasetConfig
    model_config: NCameraCNNConfig = NCameraCNNConfig()
    augmentation_config: AugmentationConfig = AugmentationConfig()
    use_train: bool = False
    device: str = "cuda"
    max_samples: Optional[int] = None


def load_checkpoint(model: NCameraCNN, path: str) -> None:
    """`model.load_state_dict(torch.load(path))` (validate.py:100-101); also accepts the `module.`-prefixed keys that
    the reference writes under --multigpu."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)


def validate(cfg: ValConfig) -> dict:
    device = torch.device(cfg.device)
    model = NCameraCNN(cfg.model_config)
    load_checkpoint(model, cfg.model_path)
    model.to(device).eval()
    dataset = CameraCubePoseDataset(cfg.dataset_config, cfg_aug=cfg.augmentation_config, train=cfg.use_train,
                                    as_uint8=True)
    losses, poses = [], []
    n = len(dataset) if cfg.max_samples is None else min(cfg.max_samples, len(dataset))
    with torch.no_grad():
        for i in range(n):
            ex = dataset[i]
            images = ex["images"].unsqueeze(0).to(device)      # (1, n_cams, H, W, 3) uint8
            target = ex["cube_pose"].unsqueeze(0).to(device)
            pred = model(images)                                # augmentation is the identity here (validate.py:106,124)
            losses.append(geometric_loss_fn(pred, target))
            poses.append(se3_exp(pred))
    losses = torch.cat(losses) if losses else torch.zeros(0)
    return {"losses": losses.cpu(), "mean_loss": float(losses.mean()) if len(losses) else float("nan"),
            "poses": torch.cat(poses).cpu() if poses else torch.zeros(0, 7)}


if __name__ == "__main__":
    import tyro

    out = validate(tyro.cli(ValConfig))
    print(f"mean loss over {len(out['losses'])} samples: {out['mean_loss']}")
