"""CameraCubePoseDataset for the reference's on-disk layout (reference: /root/reference/argus/data.py:106-229, written by
argus/data_generation.py:245-264,311-343 and by the fixture in tests/conftest.py:14-57):

    <dir>/<stem>.hdf5   attrs n_cams, H, W; groups train/ and test/ with
                        cube_poses (n,7) [x,y,z,qw,qx,qy,qz], q_leap (n,16), img_stems (n,) bytes "img/img{i}"
    <dir>/img/img{i}_{a,b}.png   H x W x 3 uint8

Metadata is read with h5py when it is importable; otherwise from the sidecar `<dir>/<stem>.npz` written by
`export_sidecar()` / `write_dataset()` (h5py is not installable in this image). Samples come back either in the
reference's format ({"images": (3*n_cams,H,W) float32 in [0,1], "cube_pose": (7,) float32 [t, q_xyzw]}, data.py:226-229)
or, for the GPU fast path, as uint8 HWC views (`as_uint8=True`): augmentation then runs on the device for the whole
batch (argus_b200.data.Augmentation) instead of per sample on CPU workers.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ROOT
from .utils import draw_spaghetti, xyzwxyz_to_xyzxyzw_SE3


@dataclass(frozen=True)
class CameraCubePoseDatasetConfig:
    """Configuration for the CameraCubePoseDataset (reference: data.py:106-142).

    Args:
        dataset_path: path to a directory holding `<name>.hdf5` (or `<name>.npz`) and `img/`.
        center_crop: (height, width) of the center crop.
    """

    dataset_path: Optional[str] = None
    center_crop: Optional[tuple[int, int]] = (256, 256)

    def __post_init__(self) -> None:
        assert isinstance(self.dataset_path, str), "The dataset path must be a str!"
        if not os.path.exists(self.dataset_path):
            if os.path.exists(ROOT + "/" + self.dataset_path):
                object.__setattr__(self, "dataset_path", ROOT + "/" + self.dataset_path)
            else:
                raise FileNotFoundError(f"The specified path does not exist: {self.dataset_path}!")
        assert not Path(self.dataset_path).suffix, "The dataset path must point to a directory!"
        if Path(self.dataset_path).is_dir():
            stem = Path(self.dataset_path).stem
            assert os.path.exists(self.dataset_path + f"/{stem}.hdf5") or os.path.exists(
                self.dataset_path + f"/{stem}.npz"), f"There must be an hdf5 (or npz sidecar) file with the name {stem}!"
            assert os.path.exists(self.dataset_path + "/img"), "The dataset must have an `img` directory!"


def _read_metadata(dataset_path: str, split: str) -> dict:
    stem = Path(dataset_path).stem
    h5 = Path(dataset_path) / f"{stem}.hdf5"
    npz = Path(dataset_path) / f"{stem}.npz"
    if h5.exists():
        try:
            import h5py  # noqa: F401
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(h5, "r") as f:
                d = f[split]
                return {"n_cams": int(f.attrs["n_cams"]), "cube_poses": d["cube_poses"][()], "q_leap": d["q_leap"][()],
                        "img_stems": [b.decode("utf-8") for b in d["img_stems"][()]]}
    if npz.exists():
        z = np.load(npz, allow_pickle=False)
        return {"n_cams": int(z["n_cams"]), "cube_poses": z[f"{split}/cube_poses"], "q_leap": z[f"{split}/q_leap"],
                "img_stems": [str(s) for s in z[f"{split}/img_stems"]]}
    raise ImportError(f"{h5} needs h5py, which is not installed; run argus_b200.dataset.export_sidecar() where h5py "
                      f"exists to create {npz}")


def export_sidecar(dataset_path: str) -> str:
    """Convert `<stem>.hdf5` into the `<stem>.npz` sidecar (run where h5py is available)."""
    import h5py

    stem = Path(dataset_path).stem
    out = {}
    with h5py.File(Path(dataset_path) / f"{stem}.hdf5", "r") as f:
        out["n_cams"] = np.int64(f.attrs["n_cams"])
        out["H"], out["W"] = np.int64(f.attrs["H"]), np.int64(f.attrs["W"])
        for split in ("train", "test"):
            out[f"{split}/cube_poses"] = f[split]["cube_poses"][()]
            out[f"{split}/q_leap"] = f[split]["q_leap"][()]
            out[f"{split}/img_stems"] = np.array([b.decode("utf-8") for b in f[split]["img_stems"][()]])
    path = str(Path(dataset_path) / f"{stem}.npz")
    np.savez(path, **out)
    return path


def write_dataset(dataset_path: str, images_train: np.ndarray, poses_train: np.ndarray, images_test: np.ndarray,
                  poses_test: np.ndarray) -> None:
    """Writes a dataset in the reference layout (PNG pairs + metadata sidecar; also the .hdf5 when h5py exists).
    images_*: (n, n_cams, H, W, 3) uint8; poses_*: (n, 7) [x, y, z, qw, qx, qy, qz] as data_generation.py stores them."""
    from PIL import Image

    root = Path(dataset_path)
    (root / "img").mkdir(parents=True, exist_ok=True)
    n_cams, H, W = images_train.shape[1:4]
    out = {"n_cams": np.int64(n_cams), "H": np.int64(H), "W": np.int64(W)}
    start = 0
    for split, imgs, poses in (("train", images_train, poses_train), ("test", images_test, poses_test)):
        stems = []
        for i in range(imgs.shape[0]):
            s = f"img/img{start + i}"
            stems.append(s)
            for v in range(n_cams):
                Image.fromarray(imgs[i, v]).save(root / f"{s}_{'abcdefgh'[v]}.png")
        start += imgs.shape[0]
        out[f"{split}/cube_poses"] = np.asarray(poses)
        out[f"{split}/q_leap"] = np.zeros((imgs.shape[0], 16))
        out[f"{split}/img_stems"] = np.array(stems)
    np.savez(root / f"{root.stem}.npz", **out)
    try:
        import h5py

        with h5py.File(root / f"{root.stem}.hdf5", "w") as f:
            f.attrs["n_cams"], f.attrs["H"], f.attrs["W"] = n_cams, H, W
            for split in ("train", "test"):
                g = f.create_group(split)
                g.create_dataset("cube_poses", data=out[f"{split}/cube_poses"])
                g.create_dataset("q_leap", data=out[f"{split}/q_leap"])
                g.create_dataset("img_stems", data=np.array([s.encode("utf-8") for s in out[f"{split}/img_stems"]]))
    except ImportError:
        pass


class CameraCubePoseDataset(Dataset):
    """The dataset for N cameras and a cube (reference: data.py:145-229)."""

    def __init__(self, cfg_dataset: CameraCubePoseDatasetConfig, cfg_aug=None, train: bool = True,
                 as_uint8: bool = False, pil_spaghetti: bool = True) -> None:
        # pil_spaghetti=False: the training loop draws the arcs on the device instead (Augmentation.spaghetti_batch)
        self.pil_spaghetti = pil_spaghetti
        meta = _read_metadata(cfg_dataset.dataset_path, "train" if train else "test")
        self.n_cams = meta["n_cams"]
        _cube_poses = torch.from_numpy(np.asarray(meta["cube_poses"]))  # stored quat order is (w, x, y, z)
        self.cube_poses = xyzwxyz_to_xyzxyzw_SE3(_cube_poses)            # (x, y, z, w), as pypose expects
        self.q_leap = torch.from_numpy(np.asarray(meta["q_leap"]))
        self.img_stems = meta["img_stems"]
        self.cfg_aug = cfg_aug
        if cfg_aug is not None:
            from .data import Augmentation

            # kept for API parity; it is applied on the device per BATCH by the training loop, not in __getitem__
            self.augmentation = Augmentation(cfg_aug, train=train)
        else:
            self.augmentation = None
        self.dataset_path = cfg_dataset.dataset_path
        self.center_crop = cfg_dataset.center_crop
        self.as_uint8 = as_uint8

    def __len__(self) -> int:
        return self.cube_poses.shape[0]

    def _load_u8(self, idx: int) -> np.ndarray:
        from PIL import Image

        stem = self.img_stems[idx]
        views = []
        for v in range(self.n_cams):
            img = Image.open(f"{self.dataset_path}/{stem}_{'abcdefgh'[v]}.png").convert("RGB")
            if self.pil_spaghetti and self.cfg_aug is not None and self.cfg_aug.num_spaghetti > 0:
                img = draw_spaghetti(img, self.cfg_aug.num_spaghetti)   # data.py:213-215 (train AND val)
            a = np.array(img)
            if self.center_crop and a.shape[:2] != tuple(self.center_crop):
                ch, cw = self.center_crop
                # kornia.center_crop offsets (data.py:219-222): floor((size - crop) / 2)
                top, left = (a.shape[0] - ch) // 2, (a.shape[1] - cw) // 2
                a = a[top:top + ch, left:left + cw]
            views.append(a)
        return np.stack(views)  # (n_cams, H, W, 3) uint8

    def __getitem__(self, idx: int) -> dict:
        u8 = self._load_u8(idx)
        pose = self.cube_poses[idx].to(torch.float32)
        if self.as_uint8:
            return {"images": torch.from_numpy(u8), "cube_pose": pose}
        images = torch.from_numpy(u8).permute(0, 3, 1, 2).reshape(-1, *u8.shape[1:3]).to(torch.float32) / 255.0
        return {"images": images, "cube_pose": pose}
