"""Validation entry point (reference: /root/reference/argus/validate.py:48-181). The compute part — load a `.pth`,
eval-mode forward at batch 1, device augmentation when `use_train`, geometric loss, pose exponential — runs on the
B200 path; the reference's matplotlib figures are host-side visualisation and out of scope:
`validate()` returns the numbers (the reference returns None after writing the figures).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import ROOT
from .data import Augmentation, AugmentationConfig
from .dataset import CameraCubePoseDataset, CameraCubePoseDatasetConfig
from .loss import geometric_loss_fn
from .models import NCameraCNN, NCameraCNNConfig
from .utils import get_tree_string, se3_exp


@dataclass(frozen=True)
class ValConfig:
    """The configuration dataclass for validation (reference: validate.py:48-82, same fields and checks).

    Fields:
        model_path: The path to the saved model to validate.
        dataset_config: The configuration for the dataset.
        model_config: The configuration for the model.
        aug_config: The configuration for the augmentation.
        use_train: Whether to use the training set (and, as in the reference, the training augmentation on device).
        device: The device to run on.
        max_samples: (not in the reference) stop after this many samples.
    """

    model_path: str
    dataset_config: CameraCubePoseDatasetConfig
    model_config: NCameraCNNConfig = NCameraCNNConfig()
    aug_config: AugmentationConfig = AugmentationConfig()
    use_train: bool = False
    device: str = "cuda" if torch.cuda.is_available() else "cpu"
    max_samples: Optional[int] = None

    def __post_init__(self) -> None:
        """Sanity checks on inputs (reference: validate.py:66-82)."""
        assert self.dataset_config is not None, (
            "The dataset config must be provided with a valid dataset path!\n"
            "Here is a tree of the `outputs/data` directory to help:\n"
            f"{get_tree_string(ROOT + '/outputs/data', 'hdf5')}"
        )
        assert isinstance(self.model_path, str), "The model path must be a str!"
        assert self.model_path.endswith(".pth"), "The model path must end with '.pth'!"
        if not os.path.exists(self.model_path):
            raise FileNotFoundError(
                f"The specified path does not exist!\n"
                f"Here is a tree of the `outputs/models` directory to help:\n"
                f"{get_tree_string(ROOT + '/outputs/models', 'pth')}"
            )


def load_checkpoint(model: NCameraCNN, path: str) -> None:
    """`model.load_state_dict(torch.load(path))` (validate.py:100-101); also accepts the `module.`-prefixed keys that
    the reference writes under --multigpu."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)


def validate(cfg: ValConfig) -> dict:
    """Validates the model on the dataset (reference: validate.py:85-181, compute part)."""
    device = torch.device(cfg.device)
    if device.type != "cuda":
        raise RuntimeError("argus_b200 validates on sm_100a GPUs only (no CPU fallback)")
    model = NCameraCNN(cfg.model_config)
    load_checkpoint(model, cfg.model_path)
    model.to(device).eval()

    # device augmentation, identity unless use_train (validate.py:106-107,124). The dataset gets cfg_aug as in the
    # reference (validate.py:110): it draws the spaghetti arcs (data.py:212-215, train AND val); its per-sample kornia
    # pass (data.py:223-225) does not exist here -- augmentation always runs on the device -- so train-set samples are
    # augmented once, below, not twice as in the reference.
    augmentation = Augmentation(cfg.aug_config, train=cfg.use_train)
    dataset = CameraCubePoseDataset(cfg.dataset_config, cfg_aug=cfg.aug_config, train=cfg.use_train, as_uint8=True)
    losses, poses = [], []
    n = len(dataset) if cfg.max_samples is None else min(cfg.max_samples, len(dataset))
    with torch.no_grad():
        for i in range(n):
            ex = dataset[i]                                      # center crop applied by the dataset (data.py:219-222)
            images = ex["images"].unsqueeze(0).to(device)       # (1, n_cams, H, W, 3) uint8
            target = ex["cube_pose"].unsqueeze(0).to(device)
            if augmentation.train and augmentation.enabled:
                images = augmentation.augment_batch(images)      # (1, 3 * n_cams, H, W) float32, augmented
            pred = model(images)
            losses.append(torch.mean(geometric_loss_fn(pred, target)).reshape(1))
            poses.append(se3_exp(pred))
    losses = torch.cat(losses) if losses else torch.zeros(0)
    return {"losses": losses.cpu(), "mean_loss": float(losses.mean()) if len(losses) else float("nan"),
            "poses": torch.cat(poses).cpu() if poses else torch.zeros(0, 7)}


if __name__ == "__main__":
    import tyro

    out = validate(tyro.cli(ValConfig))
    print(f"mean loss over {len(out['losses'])} samples: {out['mean_loss']}")
