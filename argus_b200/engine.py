"""Fused training step for the hot path of /root/reference/argus/train.py:298-320:

    images.to(device) -> [augmentation] -> model(images) -> geometric_loss_fn -> mean -> backward
    -> clip_grad_norm_(max_grad_norm) -> Adam step

executed as a handful of C-ABI calls on one stream, with no host synchronisation inside the step (the reference
synchronises every step through `loss.item()`, train.py:312). Under data parallelism (train.py:137-166,199) the
gradient arena is all-reduced in four reverse-order buckets over NCCL while the rest of the backward pass runs.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .models import NCameraCNN


def gradient_buckets(flat: torch.Tensor, stage_ranges: list[tuple[int, int]]) -> list[torch.Tensor]:
    """Views of the flat gradient arena, one per backward stage, in the order they become final
    (stage 0 = layer4 + fc + head ... stage 3 = stem + layer1). Together they tile the arena exactly once."""
    return [flat[b:e] for (b, e) in stage_ranges]


def all_reduce_bucket(bucket: torch.Tensor, group=None, async_op: bool = True):
    """SUM all-reduce of one bucket (the 1/world average is folded into the optimizer's gradient scale).
    DDP in the reference averages 25 MiB buckets the same way (train.py:199)."""
    return dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class TrainEngine:
    """Owns the optimizer state (flat Adam moments) and runs fused training steps on an NCameraCNN.

    Args:
        model: an argus_b200 NCameraCNN already on a CUDA device.
        lr, betas, eps: torch.optim.Adam defaults as used by the reference (train.py:232).
        max_grad_norm: clip_grad_norm_ threshold (train.py:318); <= 0 disables clipping.
        process_group: torch.distributed group for data parallelism (None = single process).
        augmentation: optional argus_b200.data.Augmentation applied on device to uint8 image batches.
    """

    def __init__(self, model: NCameraCNN, lr: float = 1e-4, max_grad_norm: float = 1.0,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 process_group: Optional["dist.ProcessGroup"] = None, distributed: Optional[bool] = None,
                 augmentation=None) -> None:
        if not model.flat_params.is_cuda:
            raise _lib.ArgusError("TrainEngine needs the model on a CUDA device (no CPU fallback)")
        self.model = model
        self.lr = float(lr)
        self.max_grad_norm = float(max_grad_norm)
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.augmentation = augmentation
        self.step_count = 0
        self.device = model.flat_params.device
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        self.distributed = bool(distributed)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        n = model.flat_params.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._scratch = torch.zeros(1024, dtype=torch.float32, device=self.device)
        self._grad_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._loss_mean = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._stage_ranges = model.stage_ranges()
        self.last_grad_norm = self._grad_norm
        # look-ahead staging (prefetch): side stream + events, see prefetch()
        self._side = torch.cuda.Stream(device=self.device)
        self._prev_done = torch.cuda.Event()
        self._staged_event = torch.cuda.Event()
        self._staged = None            # (data_ptr, shape) of the uint8 batch already augmented into the model
        if self.distributed:
            # DDP's constructor broadcasts rank 0's parameters and buffers (train.py:199)
            dist.broadcast(model.flat_params, src=0, group=process_group)
            dist.broadcast(model._flat_buffers, src=0, group=process_group)
            model.sync_weights(force=True)

    # ------------------------------------------------------------------------------------------------------------
    def prefetch(self, images: torch.Tensor, after: Optional["torch.cuda.Event"] = None) -> None:
        """Augment + stage the NEXT uint8 batch on a side stream while the step enqueued just before keeps the GPU
        busy (the augmentation kernel is ALU-bound, the training step HBM-bound: they overlap well). The model keeps
        two stem-input buffers, so the batch in flight is not disturbed. Call it right after step(); the following
        step() / forward_backward() must be given the same tensor. `after`: an event the staging must wait for (e.g. the
        host-to-device copy of `images` on a copy stream); work already enqueued on the current stream before the
        preceding step() is waited for automatically. A no-op for inputs the fused staging path does not cover (float
        images, fp32 mode)."""
        model = self.model
        if images.dtype != torch.uint8 or model.precision != "bf16":
            return
        aug = self.augmentation
        apply = aug is not None and aug.train and aug.enabled
        key = (images.data_ptr(), tuple(images.shape))
        # the buffer being written was last read by the step BEFORE the one in flight: wait for that step only
        self._side.wait_event(self._prev_done)
        if after is not None:
            self._side.wait_event(after)
        with torch.cuda.stream(self._side):
            if aug is not None and aug.gpu_spaghetti:
                images.record_stream(self._side)
                images = aug.spaghetti_batch(images)
            params = aug.sample_params(images.shape[0], images.shape[1], self.device) if apply else None
            B, _, H, W, _ = images.shape
            model._ensure_bound()
            with torch.cuda.device(self.device):
                _lib.call("argus_model_stage_input_u8", model._handle.ptr, images.contiguous(), params, int(B), int(H),
                          int(W), 1, int(apply), _lib.stream_ptr())
            self._staged_event.record(self._side)
        images.record_stream(self._side)
        if params is not None:
            params.record_stream(self._side)
        self._staged = key

    def forward_backward(self, images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """Forward, loss, backward (+ bucketed gradient all-reduce). Returns the mean loss as a device scalar.

        images: (B, 3*n_cams, H, W) float32 in [0,1]  or  (B, n_cams, H, W, 3) uint8; targets: (B, 7) [t, q_xyzw].
        """
        model = self.model
        lib_stream = _lib.stream_ptr()
        model.train()
        # everything enqueued so far (the previous step) precedes this event: a prefetch() issued after this call may
        # overwrite the stem buffer of the previous step as soon as it fires
        self._prev_done.record(torch.cuda.current_stream(self.device))
        staged, self._staged = self._staged, None
        if staged is not None:
            # (also when the staged batch is not the one given now: its write into the alternate buffer must be over
            # before anything else is staged there)
            torch.cuda.current_stream(self.device).wait_event(self._staged_event)
        if staged is not None and images.dtype == torch.uint8 and staged == (images.data_ptr(), tuple(images.shape)):
            out = model._forward_staged(images.shape[0], images.shape[2], images.shape[3])
        elif images.dtype == torch.uint8 and model.precision == "fp32":
            # fp32 parity mode: the same augmentation kernel with fp32 NCHW output, then the fp32 network
            aug = self.augmentation
            if aug is not None and aug.gpu_spaghetti:
                images = aug.spaghetti_batch(images)
            if aug is not None and aug.train and aug.enabled:
                out = model._forward_impl(aug.augment_batch(images), True)
            else:
                out = model._forward_impl(images, True)
        elif images.dtype == torch.uint8:
            # fused augmentation + staging: uint8 pairs -> augmented bf16 stem input inside the model's arena
            aug = self.augmentation
            if aug is not None and aug.gpu_spaghetti:
                images = aug.spaghetti_batch(images)
            apply = aug is not None and aug.train and aug.enabled
            params = aug.sample_params(images.shape[0], images.shape[1], self.device) if apply else None
            out = model._forward_impl(images, True, aug_params=params, augment=apply)
        else:
            out = model._forward_impl(images, True)
        B = out.shape[0]
        targets = targets.to(device=self.device, dtype=torch.float32).contiguous()
        grad = torch.empty_like(out)
        self._loss_mean.zero_()
        with torch.cuda.device(self.device):
            # d(mean loss)/d pred = grad / B   (train.py:308-309; the loss is always evaluated in fp32)
            _lib.call("argus_pose_loss", out, targets, None, self._loss_mean, grad, int(B), 1.0 / B, lib_stream)
            _lib.call("argus_model_zero_grads", model._handle.ptr, lib_stream)
            works = []
            buckets = gradient_buckets(model.flat_grads, self._stage_ranges)
            for stage in range(4):
                _lib.call("argus_model_backward", model._handle.ptr, grad, stage, stage + 1, lib_stream)
                if self.world > 1:
                    # NCCL runs on its own stream: it waits for this stage's kernels, the next stage overlaps with it
                    works.append(all_reduce_bucket(buckets[stage], self.group, async_op=True))
            for w in works:
                w.wait()
        return self._loss_mean[0]

    def optimizer_step(self) -> None:
        """clip_grad_norm_ + Adam on the flat arenas, then refresh the packed bf16 weights."""
        model = self.model
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.argus_clip_adam_step(
                _lib.ptr(model.flat_params), _lib.ptr(model.flat_grads), _lib.ptr(self.exp_avg),
                _lib.ptr(self.exp_avg_sq), ctypes.c_int64(model.flat_params.numel()), _lib.ptr(self._scratch),
                ctypes.c_float(1.0 / self.world), ctypes.c_float(self.max_grad_norm), ctypes.c_float(self.lr),
                ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]), ctypes.c_float(self.eps),
                ctypes.c_int(self.step_count), _lib.ptr(self._grad_norm), _lib.stream_ptr()))
        model.sync_weights(force=True)

    def step(self, images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """One full training step; returns the (pre-update) mean loss as a device scalar without synchronising."""
        loss = self.forward_backward(images, targets)
        self.optimizer_step()
        return loss

    # ------------------------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr}

    def load_state_dict(self, state: dict) -> None:
        self.step_count = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.lr = float(state["lr"])
