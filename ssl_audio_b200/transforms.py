"""GPU mirror of the reference's two-view transform (reference: utils/transforms.py:7-58).

`AudioPairTransform(args, ...)` keeps the reference signature and reads the same `args` fields
(mixup, Gnoise, RRC, RLF, n_mels, crop_frames, virtual_crop_scale, local_crops_number,
local_crops_size).  `forward(x)` takes a normalised log-mel `(1, F, T)` -- or a batch
`(B, 1, F, T)`, processed in the reference's sequential sample order -- and returns
`[global(x), global(x)] + [local(x)] * local_crops_number`; for a batch each entry is the
collated `(B, 1, ., .)` tensor the reference's DataLoader would have produced.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .augmentations import ViewEngine, _as_batch, _check_cuda_f32, push_bank, run_views
from .planner import ViewPlanner

__all__ = ["AudioPairTransform"]


class AudioPairTransform(nn.Module):

    def __init__(self, args, train_transform=True, multi_transform=True,
                 mixup_ratio=0.2, gauss_noise_ratio=0.2,
                 global_crop_scale=(0.6, 1.5), local_crop_scale=(0.05, 0.6)):
        super().__init__()
        self.multi_transform = multi_transform
        self.local_crops_number = args.local_crops_number
        self.train_transform = train_transform
        self._gnoise = bool(train_transform is True and getattr(args, "Gnoise", False))
        self._gnoise_ratio = gauss_noise_ratio
        self._mixup = bool(train_transform is True and args.mixup)
        self._rrc = bool(train_transform is True and args.RRC)
        self._rlf = bool(train_transform is True and args.RLF)
        self._in_hw = (int(args.n_mels), int(args.crop_frames))
        vcs = tuple(args.virtual_crop_scale)
        self._canvas_hw = (int(self._in_hw[0] * vcs[0]), int(self._in_hw[1] * vcs[1])) if self._rrc else self._in_hw
        self._local_hw = tuple(int(v) for v in args.local_crops_size)
        self._mixup_ratio = mixup_ratio
        self._global_crop_scale = tuple(global_crop_scale)
        self._local_crop_scale = tuple(local_crop_scale)
        self._n_global = 2 if multi_transform else 1
        self._n_local = int(self.local_crops_number) if multi_transform else 0
        self._engine = None
        self._max_batch = 1024
        # The reference's stage-by-stage pipelines (utils/transforms.py:15-46), for code that uses them directly or prints them:
        # nn.Sequential of the standalone GPU stages, nn.Identity without train_transform.  `forward` does not go through them (it
        # runs all stages of all views in one kernel, with its own Mixup ring); a single-view call `global_transform(x)` draws from
        # the same global generators in the reference's order and keeps its own memory bank, as the reference's MixupBYOLA does.
        from . import augmentations as A
        if train_transform is True:
            stages = []
            if args.mixup:
                stages.append(A.MixupBYOLA(ratio=mixup_ratio))
            if getattr(args, "Gnoise", False):
                stages.append(A.MixGaussianNoise(ratio=gauss_noise_ratio))
            if args.RRC:
                stages.append(A.RandomResizeCrop((args.n_mels, args.crop_frames), virtual_crop_scale=tuple(args.virtual_crop_scale),
                                                 freq_scale=global_crop_scale, time_scale=global_crop_scale))
            if args.RLF:
                stages.append(A.RandomLinearFader())
            self.global_transform = nn.Sequential(*stages)
        else:
            self.global_transform = nn.Identity()
        self.local_transform = nn.Sequential(A.RandomResizeCrop(tuple(args.local_crops_size), virtual_crop_scale=(1, 1),
                                                                freq_scale=local_crop_scale, time_scale=local_crop_scale))

    # ------------------------------------------------------------------------------------------
    def engine(self, batch: int) -> ViewEngine:
        """The planner + device ring, (re)built when a larger batch needs a larger ring."""
        if self._engine is None or batch > self._max_batch:
            if self._engine is not None and self._engine.planner.bank_len() > 0:
                raise RuntimeError(f"batch of {batch} exceeds the Mixup ring sized for {self._max_batch}; "
                                   "construct AudioPairTransform and call .reserve(batch) before the first forward")
            self._max_batch = max(self._max_batch, batch)
            pl = ViewPlanner(mixup=self._mixup, rrc=self._rrc, rlf=self._rlf, mixup_ratio=self._mixup_ratio, n_memory=2048,
                             ring_slots=2048 + self._max_batch, n_global=self._n_global, in_hw=self._in_hw,
                             canvas_hw=self._canvas_hw, freq_scale=self._global_crop_scale, time_scale=self._global_crop_scale,
                             n_local=self._n_local, local_hw=self._local_hw, local_scale=self._local_crop_scale,
                             gnoise=self._gnoise, gnoise_ratio=self._gnoise_ratio)
            self._engine = ViewEngine(pl, self._in_hw, self._canvas_hw)
        return self._engine

    def reserve(self, batch: int) -> None:
        self.engine(int(batch))

    @property
    def memory_bank_len(self) -> int:
        return 0 if self._engine is None else self._engine.planner.bank_len()

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """`noise` (only with args.Gnoise): the N(0, 1) draws of MixGaussianNoise, (B, n_global, F, T) -- see MixGaussianNoise."""
        _check_cuda_f32(x, "x")
        x4, single = _as_batch(x.contiguous())
        if (int(x4.shape[2]), int(x4.shape[3])) != self._in_hw:
            raise ValueError(f"expected log-mel of shape {self._in_hw}, got {tuple(x4.shape[2:])}")
        B = int(x4.shape[0])
        eng = self.engine(B)
        if self._mixup:
            eng.ensure_ring(x4.device)
        plan = eng.planner.plan(B, device=x4.device)
        outs = self.views_from_plan(x4, 0, self._in_hw[0] * self._in_hw[1], plan, noise)
        if self._mixup:
            push_bank(eng, x4, self._in_hw[0] * self._in_hw[1], plan)
        if single:
            outs = [o[0] for o in outs]
        return outs if self.multi_transform else outs[0]

    def views_from_plan(self, x, x_slot_ptr, x_slot_stride, plan, noise=None) -> List[torch.Tensor]:
        """Views of an already planned batch whose clips live at x + x_slot[b] * x_slot_stride (x_slot_ptr: device pointer to
        int32 slots, 0 = identity; used by the batch frontend, which lets the log-mel kernel write the clips straight into
        the Mixup ring)."""
        return run_views(self._engine, x, x_slot_ptr, x_slot_stride, plan, self._n_global, self._n_local, self._in_hw, self._local_hw,
                         noise)
