"""Data path of the hot loop (reference: /root/reference/argus/data.py).

* AugmentationConfig / Augmentation keep the reference's fields and call convention (data.py:18-103) but run as one
  fused CUDA kernel on the whole batch instead of kornia ops on CPU inside Dataset.__getitem__.
* CameraCubePoseDataset (data.py:145-229) is provided by argus_b200.dataset (reader for the reference's on-disk
  layout without h5py); it is re-exported here under the reference's names.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Union

import torch

from . import _lib


@dataclass(frozen=True)
class AugmentationConfig:
    """Configuration for data augmentation (same fields and defaults as the reference, data.py:18-38)."""

    # color jiggle
    brightness: Union[float, tuple[float, float]] = (0.8, 1.0)
    contrast: Union[float, tuple[float, float]] = (0.5, 1.2)
    saturation: Union[float, tuple[float, float]] = (0.25, 1.2)
    hue: Union[float, tuple[float, float]] = (-0.1, 0.1)

    # spaghetti
    num_spaghetti: int = 10

    # flags
    color_jiggle: bool = True
    planckian_jitter: bool = True
    random_erasing: bool = False
    blur: bool = True
    motion_blur: bool = True
    plasma_shadow: bool = True
    salt_and_pepper: bool = False


class _AugConfigC(ctypes.Structure):
    _fields_ = [("color_jiggle", ctypes.c_int), ("planckian_jitter", ctypes.c_int), ("blur", ctypes.c_int),
                ("motion_blur", ctypes.c_int), ("plasma_shadow", ctypes.c_int),
                ("brightness_lo", ctypes.c_float), ("brightness_span", ctypes.c_float),
                ("contrast_lo", ctypes.c_float), ("contrast_span", ctypes.c_float),
                ("saturation_lo", ctypes.c_float), ("saturation_span", ctypes.c_float),
                ("hue_lo", ctypes.c_float), ("hue_span", ctypes.c_float),
                ("random_erasing", ctypes.c_int), ("salt_and_pepper", ctypes.c_int)]


def _range(v, center: float, lower_bound: Optional[float] = None) -> tuple[float, float]:
    """kornia's convention for a scalar factor f: the range is (center - f, center + f)."""
    if isinstance(v, (tuple, list)):
        return float(v[0]), float(v[1])
    lo, hi = center - float(v), center + float(v)
    if lower_bound is not None:
        lo = max(lo, lower_bound)
    return lo, hi


N_PARAMS = 40     # ARGUS_AUG_PARAMS
ARC_FIELDS = 8    # ARGUS_ARC_FIELDS


class Augmentation(torch.nn.Module):
    """Data augmentation module for the images (reference: data.py:41-103).

    forward(images) keeps the reference contract — float images (n, 3, H, W) in [0,1] in, same shape out, identity
    unless constructed with train=True — and treats the n images as the n_cams views of ONE sample, exactly as the
    reference does when it calls the module per sample (data.py:224). `augment_batch` is the batched entry point used
    by the training engine: uint8 (B, n_cams, H, W, 3) pairs -> augmented images, one launch for the whole batch.

    Every stage of the reference's kornia chain exists, in its order (data.py:52-95): RandomErasing x2 and
    RandomSaltAndPepperNoise (default-off flags), RandomPlanckianJitter, ColorJiggle (same_on_batch), RandomGaussianBlur,
    RandomMotionBlur, RandomPlasmaShadow (kornia's diamond-square fractal). The arithmetic spec is oracle/augment.py.

    Randomness: every sampled parameter is a pure function of (seed, step, image index), see oracle/augment.py; the
    module-level counter `step` advances on every call so that successive calls differ and a re-run with the same
    seed reproduces them (the reference's only augmentation contract, tests/test_train.py:69-77).
    """

    def __init__(self, cfg: AugmentationConfig, train: bool = True, seed: Optional[int] = None,
                 gpu_spaghetti: bool = False) -> None:
        super().__init__()
        self.cfg = cfg
        self.train = train  # (shadows nn.Module.train like the reference does, data.py:48)
        self.seed = int(torch.initial_seed() if seed is None else seed) & ((1 << 63) - 1)
        self.step = 0
        b, c = _range(cfg.brightness, 1.0, 0.0), _range(cfg.contrast, 1.0, 0.0)
        s, h = _range(cfg.saturation, 1.0, 0.0), _range(cfg.hue, 0.0)
        self._c = _AugConfigC(int(cfg.color_jiggle), int(cfg.planckian_jitter), int(cfg.blur), int(cfg.motion_blur),
                              int(cfg.plasma_shadow), b[0], b[1] - b[0], c[0], c[1] - c[0], s[0], s[1] - s[0],
                              h[0], h[1] - h[0], int(cfg.random_erasing), int(cfg.salt_and_pepper))
        self.enabled = any([cfg.color_jiggle, cfg.planckian_jitter, cfg.blur, cfg.motion_blur, cfg.plasma_shadow,
                            cfg.random_erasing, cfg.salt_and_pepper])
        # draw_spaghetti (reference: utils.py:252-275, applied by the dataset with PIL at data.py:212-215, train AND
        # val). gpu_spaghetti=True moves it onto the device (spaghetti_batch / arc_mask, the same pixels Pillow paints);
        # the dataset must then skip its PIL pass.
        self.gpu_spaghetti = bool(gpu_spaghetti) and cfg.num_spaghetti > 0
        self.spaghetti_step = 0
        self._ws: dict = {}      # (device, n_images, H, W, tag) -> int32 bit-mask workspace

    # ------------------------------------------------------------------------------------------------------------
    def workspace(self, n_images: int, H: int, W: int, device, tag: str = "plasma") -> torch.Tensor:
        """int32 [n_images, H, W // 32] bit-mask buffer (plasma shadow mask / arc mask), cached per shape."""
        key = (str(device), n_images, H, W, tag)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) > 8:
                self._ws.clear()
            ws = torch.empty((n_images, H, W // 32), dtype=torch.int32, device=device)
            self._ws[key] = ws
        return ws

    def sample_params(self, n_pairs: int, n_cams: int, device, step: Optional[int] = None, H: int = 256,
                      W: int = 256) -> torch.Tensor:
        """(n_pairs*n_cams, N_PARAMS) fp32 parameter table on `device` for the given step (default: internal counter).
        H, W: image size (the random-erasing rectangles are sampled in pixels)."""
        if step is None:
            step = self.step
            self.step += 1
        params = torch.empty((n_pairs * n_cams, N_PARAMS), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.load().argus_augment_sample_params(
                _lib.ptr(params), ctypes.c_int(n_pairs * n_cams), ctypes.c_int(n_cams), ctypes.c_int(H), ctypes.c_int(W),
                ctypes.c_uint64(self.seed), ctypes.c_uint64(int(step)), ctypes.byref(self._c), _lib.stream_ptr()))
        return params

    def arc_params(self, n_images: int, H: int, W: int, device, step: Optional[int] = None) -> torch.Tensor:
        """(n_images, num_spaghetti, ARC_FIELDS) fp32 arc table [x0, y0, x1, y1, start, end, width, 0], sampled as
        draw_spaghetti does (utils.py:265-270); a pure function of (seed, step, image, arc)."""
        if step is None:
            step = self.spaghetti_step
            self.spaghetti_step += 1
        n_arcs = int(self.cfg.num_spaghetti)
        arcs = torch.empty((n_images, n_arcs, ARC_FIELDS), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            # a different stream of the hash than the augmentation parameters: fields 1000+ (oracle/augment.py)
            _lib.check(_lib.load().argus_spaghetti_sample_params(
                _lib.ptr(arcs), ctypes.c_int(n_images), ctypes.c_int(n_arcs), ctypes.c_int(H), ctypes.c_int(W),
                ctypes.c_uint64(self.seed), ctypes.c_uint64(int(step)), _lib.stream_ptr()))
        return arcs

    def arc_mask(self, n_images: int, H: int, W: int, device, step: Optional[int] = None,
                 arcs: Optional[torch.Tensor] = None, tag: str = "arcs") -> torch.Tensor:
        """int32 [n_images, H, W // 32]: bit x % 32 of word (n, y, x // 32) is set where an arc covers pixel (y, x)."""
        if arcs is None:
            arcs = self.arc_params(n_images, H, W, device, step)
        mask = self.workspace(n_images, H, W, device, tag=tag)
        with torch.cuda.device(device):
            _lib.call("argus_spaghetti_mask", arcs, mask, int(n_images), int(arcs.shape[1]), int(H), int(W),
                      _lib.stream_ptr())
        return mask

    def spaghetti_batch(self, images: torch.Tensor, step: Optional[int] = None) -> torch.Tensor:
        """uint8 (B, n_cams, H, W, 3) -> a copy with `num_spaghetti` random black arcs per image (the pixels Pillow's
        ImageDraw.arc paints, argus_spaghetti_*). Arcs are a pure function of (seed, step, image)."""
        if not images.is_cuda or images.dtype != torch.uint8:
            raise _lib.ArgusError("spaghetti_batch takes uint8 CUDA images (no CPU fallback; the PIL path is "
                                  "argus_b200.utils.draw_spaghetti)")
        B, n_cams, H, W, _ = images.shape
        images = images.contiguous()
        arcs = self.arc_params(B * n_cams, H, W, images.device, step)
        ws = self.workspace(B * n_cams, H, W, images.device, tag="arcs")
        out = torch.empty_like(images)
        with torch.cuda.device(images.device):
            _lib.call("argus_spaghetti_draw", images, out, arcs, ws, int(B * n_cams), int(arcs.shape[1]), int(H), int(W),
                      _lib.stream_ptr())
        return out

    def augment_batch(self, images: torch.Tensor, params: Optional[torch.Tensor] = None, step: Optional[int] = None,
                      arc_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 (B, n_cams, H, W, 3) or float (B, 3*n_cams, H, W) -> float32 (B, 3*n_cams, H, W) augmented.
        arc_mask: optional spaghetti bit mask (uint8 input), painted before everything else."""
        if not images.is_cuda:
            raise _lib.ArgusError("Augmentation runs on sm_100a GPUs only (no CPU fallback)")
        if images.dtype == torch.uint8:
            B, n_cams, H, W, _ = images.shape
        else:
            B, C, H, W = images.shape
            n_cams = C // 3
            images = images.to(torch.float32)
        images = images.contiguous()
        apply = bool(self.train and self.enabled)
        if params is None and apply:
            params = self.sample_params(B, n_cams, images.device, step, H=H, W=W)
        ws = self.workspace(B * n_cams, H, W, images.device) if apply else None
        out = torch.empty((B, 3 * n_cams, H, W), dtype=torch.float32, device=images.device)
        with torch.cuda.device(images.device):
            _lib.call("argus_augment", images, int(images.dtype == torch.uint8), out, 0, params, arc_mask, ws,
                      int(B * n_cams), int(H), int(W), int(apply), _lib.stream_ptr())
        return out

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """Applies the augmentations to the (n_cams, 3, H, W) views of one sample (reference: data.py:99-103)."""
        if not (self.enabled and self.train):
            return images
        n, c, H, W = images.shape
        assert c == 3, "Augmentation.forward expects (n_cams, 3, H, W) images"
        out = self.augment_batch(images.reshape(1, n * 3, H, W))
        return out.reshape(n, 3, H, W).to(images.dtype)


def __getattr__(name):
    # lazy re-export so that `from argus_b200.data import CameraCubePoseDataset` works like the reference module
    if name in ("CameraCubePoseDataset", "CameraCubePoseDatasetConfig"):
        from . import dataset

        return getattr(dataset, name)
    raise AttributeError(name)
