"""Raw-shard dataset writer and the Python face of the native double-buffered loader (csrc/loader.cu).

The reference streams samples through `torch.utils.data.DataLoader` workers that decode two PNGs per sample and
augment on the CPU (argus/train.py:147-192, argus/data.py:206-229). For large datasets (BASELINE.json configs[4]) the
B200 path pre-decodes once into a raw uint8 shard (`convert_dataset`) and then streams batches with one background
thread, pinned double buffers and a side-stream H2D copy; augmentation happens on the GPU per batch.
"""
from __future__ import annotations

import ctypes
import struct
from pathlib import Path
from typing import Iterator, Optional

import numpy as np
import torch

from . import _lib

HEADER = struct.Struct("<8sIIIIQQQ16x")
MAGIC = b"ARGUSRAW"


def write_shard(path: str, images: np.ndarray, poses_xyzw: np.ndarray) -> None:
    """images (n, n_cams, H, W, 3) uint8, poses (n, 7) [x, y, z, qx, qy, qz, qw] -> ARGUSRAW v1 file."""
    n, n_cams, H, W, c = images.shape
    assert c == 3 and images.dtype == np.uint8 and poses_xyzw.shape == (n, 7)
    pose_off = HEADER.size
    img_off = (pose_off + n * 28 + 4095) // 4096 * 4096
    with open(path, "wb") as f:
        f.write(HEADER.pack(MAGIC, 1, n_cams, H, W, n, pose_off, img_off))
        f.write(np.ascontiguousarray(poses_xyzw, dtype=np.float32).tobytes())
        f.write(b"\0" * (img_off - pose_off - n * 28))
        f.write(np.ascontiguousarray(images).tobytes())


def write_synthetic_shard(path: str, n: int, n_cams: int = 2, H: int = 256, W: int = 256, seed: int = 0,
                          chunk: int = 512, fast: bool = False) -> None:
    """Synthetic dataset of the reference's shape (uniform-noise images as in tests/conftest.py:35-41, random SE3
    poses), written chunk by chunk so that it scales to the medium / large split sizes. fast=True draws ONE random
    chunk and writes it rotated by the chunk index (I/O-bound instead of RNG-bound: a 51 GB "medium" shard in a minute;
    every sample is still uniform noise, samples repeat every `chunk`)."""
    rng = np.random.default_rng(seed)
    pose_off = HEADER.size
    img_off = (pose_off + n * 28 + 4095) // 4096 * 4096
    with open(path, "wb") as f:
        f.write(HEADER.pack(MAGIC, 1, n_cams, H, W, n, pose_off, img_off))
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        f.write(np.concatenate([rng.normal(size=(n, 3)), q], -1).astype(np.float32).tobytes())
        f.write(b"\0" * (img_off - pose_off - n * 28))
        base = rng.integers(0, 256, (chunk, n_cams, H, W, 3), dtype=np.uint8) if fast else None
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            if fast:
                k = (lo // chunk) % chunk
                f.write(memoryview(np.ascontiguousarray(np.concatenate([base[k:], base[:k]])[:m])))
            else:
                f.write(rng.integers(0, 256, (m, n_cams, H, W, 3), dtype=np.uint8).tobytes())


def convert_dataset(dataset_path: str, split: str, out_path: str, center_crop=(256, 256)) -> None:
    """Pre-decode a reference-layout dataset (hdf5/npz + PNGs) into a raw shard."""
    from .dataset import CameraCubePoseDataset, CameraCubePoseDatasetConfig

    ds = CameraCubePoseDataset(CameraCubePoseDatasetConfig(dataset_path=dataset_path, center_crop=center_crop),
                               cfg_aug=None, train=(split == "train"), as_uint8=True)
    imgs = np.stack([ds[i]["images"].numpy() for i in range(len(ds))])
    write_shard(out_path, imgs, ds.cube_poses.numpy().astype(np.float32))


class ShardLoader:
    """Iterates (images uint8 (b, n_cams, H, W, 3), poses float32 (b, 7)) device tensors over the rank's slice of a shard.

    Two device buffers: the tensors of batch k are overwritten by the fetch of batch k + 2, which waits (on the device)
    for the work that read them. `lookahead=False` (default) is for fetch-then-step loops (`for x, t in loader:
    engine.step(x, t)`): the work on batch k is whatever was enqueued before the fetch of batch k + 1.
    `lookahead=True` is for loops that fetch batch k + 1 before they enqueue step k (TrainEngine.prefetch pipelines):
    the work on batch k is whatever was enqueued before the fetch of batch k + 2."""

    def __init__(self, path: str, batch_size: int, device, rank: int = 0, world: int = 1, seed: int = 0,
                 shuffle: bool = True, drop_last: bool = False, lookahead: bool = False) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ArgusError("ShardLoader stages batches on a CUDA device")
        lib = _lib.load()
        self._ptr = ctypes.c_void_p()
        _lib.check(lib.argus_loader_create(ctypes.byref(self._ptr), str(Path(path)).encode(), ctypes.c_int(batch_size),
                                           ctypes.c_int(rank), ctypes.c_int(world), ctypes.c_uint64(seed),
                                           ctypes.c_int(int(shuffle)), ctypes.c_int(int(drop_last))))
        n, nc, H, W = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        per, nb = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(lib.argus_loader_info(self._ptr, ctypes.byref(n), ctypes.byref(nc), ctypes.byref(H), ctypes.byref(W),
                                         ctypes.byref(per), ctypes.byref(nb)))
        self.n_samples, self.n_cams, self.H, self.W = n.value, nc.value, H.value, W.value
        self.samples_per_rank, self.batches_per_epoch = per.value, nb.value
        self.batch_size = batch_size
        shape = (batch_size, self.n_cams, self.H, self.W, 3)
        self._host_img = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._host_pose = [torch.empty((batch_size, 7), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._dev_img = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._dev_pose = [torch.empty((batch_size, 7), dtype=torch.float32, device=self.device) for _ in range(2)]
        with torch.cuda.device(self.device):
            _lib.check(lib.argus_loader_bind(self._ptr, *[_lib.ptr(t) for t in self._host_img],
                                             *[_lib.ptr(t) for t in self._host_pose], *[_lib.ptr(t) for t in self._dev_img],
                                             *[_lib.ptr(t) for t in self._dev_pose]))
        _lib.check(lib.argus_loader_set_lookahead(self._ptr, ctypes.c_int(int(lookahead))))
        self.epoch = 0

    def __len__(self) -> int:
        return self.batches_per_epoch

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def __iter__(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.argus_loader_start_epoch(self._ptr, ctypes.c_int(self.epoch)))
        count, buf = ctypes.c_int(), ctypes.c_int()
        while True:
            with torch.cuda.device(self.device):
                _lib.check(lib.argus_loader_next(self._ptr, _lib.stream_ptr(), ctypes.byref(count), ctypes.byref(buf)))
            if count.value == 0:
                break
            yield self._dev_img[buf.value][: count.value], self._dev_pose[buf.value][: count.value]

    def close(self) -> None:
        if getattr(self, "_ptr", None):
            _lib.load().argus_loader_destroy(self._ptr)
            self._ptr = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
