"""GPU mirror of the reference's log-mel augmentations (reference: augmentations.py).

Same class names, constructor arguments and `forward` meaning as the reference; every `forward`
additionally accepts a batch `(B, 1, F, T)` and then returns the collate layout `(B, 1, F, T)`.
Inputs must be CUDA fp32 tensors; all arithmetic runs in the hand-written view kernel
(csrc/frontend.cu: views_kernel) and all random draws go through the native host planner, which
replays numpy's and CPython's global generators in the reference's call order.  There is no CPU
path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .planner import VIEW_DTYPE, BatchPlan, ViewPlanner

__all__ = ["RandomResizeCrop", "RandomLinearFader", "MixupBYOLA", "log_mixup_exp", "MixGaussianNoise", "NormalizeBatch", "RunningNorm", "ViewEngine"]


def _as_batch(x: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """(1,F,T) -> (1,1,F,T) view; (B,1,F,T) unchanged.  Returns (batch, was_single)."""
    if x.dim() == 3:
        if x.shape[0] != 1:
            raise ValueError(f"expected a single-channel log-mel (1, F, T), got {tuple(x.shape)}")
        return x.unsqueeze(0), True
    if x.dim() == 4 and x.shape[1] == 1:
        return x, False
    raise ValueError(f"expected (1, F, T) or (B, 1, F, T), got {tuple(x.shape)}")


def _check_cuda_f32(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: ssl_audio_b200 has no CPU path")
    if x.dtype != torch.float32:
        raise ValueError(f"{what} must be float32 (the reference forces .to(torch.float), augmentations.py:55,117)")


class ViewEngine:
    """Owns the device Mixup ring and launches the view kernel for a planned batch."""

    def __init__(self, planner: ViewPlanner, in_hw: Tuple[int, int], canvas_hw: Tuple[int, int]):
        self.planner = planner
        self.in_hw = tuple(int(v) for v in in_hw)
        self.canvas_hw = tuple(int(v) for v in canvas_hw)
        self.ring: Optional[torch.Tensor] = None      # (ring_slots, F*T) fp32
        self._lib = _lib.load()

    def ensure_ring(self, device: torch.device) -> torch.Tensor:
        if self.ring is None or self.ring.device != device:
            # rows are only ever read after they were written (the planner never hands out an unwritten slot)
            self.ring = torch.empty((self.planner.ring_slots, self.in_hw[0] * self.in_hw[1]), dtype=torch.float32, device=device)
        return self.ring

    @staticmethod
    def upload(arr: np.ndarray, device: torch.device) -> torch.Tensor:
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
        return t.to(device, non_blocking=True)


def _launch_views(lib, x_ptr: int, x_slot_ptr: int, x_slot_stride: int, ring: Optional[torch.Tensor], params_ptr: int, param_stride: int,
                  view_offset: int, n_clips: int, n_views: int, in_hw, canvas_hw, out_hw, outs: List[torch.Tensor], device,
                  noise: Optional[torch.Tensor] = None) -> None:
    a = _lib.ViewsArgs()
    a.n_clips, a.n_views = int(n_clips), int(n_views)
    a.in_h, a.in_w = int(in_hw[0]), int(in_hw[1])
    a.canvas_h, a.canvas_w = int(canvas_hw[0]), int(canvas_hw[1])
    a.out_h, a.out_w = int(out_hw[0]), int(out_hw[1])
    a.param_stride, a.view_offset = int(param_stride), int(view_offset)
    a.x = x_ptr
    a.x_slot = x_slot_ptr if x_slot_ptr else None
    a.x_slot_stride = int(x_slot_stride)
    a.bank = ring.data_ptr() if ring is not None else None
    a.bank_slot_stride = int(ring.shape[1]) if ring is not None else 0
    a.params = params_ptr
    for k, o in enumerate(outs):
        a.outs[k] = o.data_ptr()
    if noise is not None:
        a.noise, a.noise_views = noise.data_ptr(), int(noise.shape[1])
    _lib.check(lib.abt_views_fwd(C.byref(a), torch.cuda.current_stream(device).cuda_stream))


def run_views(engine: ViewEngine, x: torch.Tensor, x_slot_ptr: int, x_slot_stride: int, plan: BatchPlan,
              n_global: int, n_local: int, global_out_hw, local_out_hw, noise: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """Run all views of a planned (and uploaded) batch: one launch for the global views, one for the local crops.
    `noise` (MixGaussianNoise only): standard-normal draws (n_clips, n_global, F, T); drawn on the device when not given."""
    lib = engine._lib
    dev = x.device
    n_clips = plan.params.shape[0]
    n_views = n_global + n_local
    if engine.planner.gnoise and n_global:
        shape = (n_clips, n_global, engine.in_hw[0], engine.in_hw[1])
        if noise is None:
            # the reference draws torch.normal(0, lambd, shape) from torch's CPU generator (augmentations.py:137); here the N(0, 1)
            # draws come from torch's CUDA generator (seeded by torch.manual_seed) and the kernel scales them by lambd
            noise = torch.randn(shape, dtype=torch.float32, device=dev)
        else:
            _check_cuda_f32(noise, "noise")
            if tuple(noise.shape) != shape or not noise.is_contiguous():
                raise ValueError(f"noise must be contiguous with shape {shape}, got {tuple(noise.shape)}")
    else:
        noise = None
    outs: List[torch.Tensor] = []
    if n_global:
        g_outs = [torch.empty((n_clips, 1, global_out_hw[0], global_out_hw[1]), dtype=torch.float32, device=dev) for _ in range(n_global)]
        if n_clips:
            _launch_views(lib, x.data_ptr(), x_slot_ptr, x_slot_stride, engine.ring, plan.params_ptr, n_views, 0, n_clips, n_global,
                          engine.in_hw, engine.canvas_hw, global_out_hw, g_outs, dev, noise)
        outs += g_outs
    if n_local:
        l_outs = [torch.empty((n_clips, 1, local_out_hw[0], local_out_hw[1]), dtype=torch.float32, device=dev) for _ in range(n_local)]
        if n_clips:
            # local crops use virtual_crop_scale (1, 1): the canvas is the input itself
            _launch_views(lib, x.data_ptr(), x_slot_ptr, x_slot_stride, None, plan.params_ptr, n_views, n_global, n_clips, n_local,
                          engine.in_hw, engine.in_hw, local_out_hw, l_outs, dev)
        outs += l_outs
    return outs


def push_bank(engine: ViewEngine, x: torch.Tensor, x_stride: int, plan: BatchPlan) -> None:
    """bank[slot[b]] = x[b], stream-ordered AFTER the view kernel that may still read the old slot contents."""
    n = int(plan.slots.shape[0])
    if n == 0:
        return
    ring = engine.ring
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(engine._lib.abt_bank_push(x.data_ptr(), int(x_stride), n, int(ring.shape[1]), ring.data_ptr(), int(ring.shape[1]),
                                         plan.slots_ptr, stream))


class _SingleStage(nn.Module):
    """Base of the three stand-alone augmentation modules (they share the view kernel and the planner)."""


class RandomResizeCrop(_SingleStage):
    """Random Resize Crop block (reference: augmentations.py:12-61).

    Args mirror the reference: out_size, virtual_crop_scale `(F ratio, T ratio)`, freq_scale, time_scale.
    """

    def __init__(self, out_size=(64, 96), virtual_crop_scale=(1.0, 1.5), freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5)):
        super().__init__()
        self.out_size = out_size
        self.virtual_crop_scale = virtual_crop_scale
        self.freq_scale = freq_scale
        self.time_scale = time_scale
        self.interpolation = "bicubic"
        self._engine = None

    @staticmethod
    def get_params(virtual_crop_size, in_size, time_scale, freq_scale):
        """Same draws as the reference (augmentations.py:30-38), through the native planner.  Note the
        reference passes (time_scale, freq_scale) positionally: h is drawn from `freq_scale`, w from `time_scale`."""
        canvas_h, canvas_w = (int(v) for v in virtual_crop_size)
        src_h, src_w = (int(v) for v in in_size)
        pl = ViewPlanner(mixup=False, rrc=True, rlf=False, n_global=1, in_hw=(src_h, src_w), canvas_hw=(canvas_h, canvas_w),
                         freq_scale=freq_scale, time_scale=time_scale)
        p = pl.plan(1).params[0, 0]
        return int(p["i"]), int(p["j"]), int(p["h"]), int(p["w"])

    def forward(self, lms: torch.Tensor) -> torch.Tensor:
        _check_cuda_f32(lms, "lms")
        x4, single = _as_batch(lms.contiguous())
        F, T = int(x4.shape[2]), int(x4.shape[3])
        canvas = (int(F * self.virtual_crop_scale[0]), int(T * self.virtual_crop_scale[1]))
        if self._engine is None or self._engine.in_hw != (F, T):
            pl = ViewPlanner(mixup=False, rrc=True, rlf=False, n_global=1, in_hw=(F, T), canvas_hw=canvas,
                             freq_scale=self.freq_scale, time_scale=self.time_scale)
            self._engine = ViewEngine(pl, (F, T), canvas)
        plan = self._engine.planner.plan(x4.shape[0], device=x4.device)
        out = run_views(self._engine, x4, 0, F * T, plan, 1, 0, tuple(self.out_size), None)[0]
        return out[0] if single else out

    def __repr__(self):
        format_string = self.__class__.__name__ + f"(virtual_crop_size={self.virtual_crop_scale}"
        format_string += ", time_scale={0}".format(tuple(round(s, 4) for s in self.time_scale))
        format_string += ", freq_scale={0})".format(tuple(round(r, 4) for r in self.freq_scale))
        return format_string


class RandomLinearFader(_SingleStage):
    """reference: augmentations.py:64-78."""

    def __init__(self, gain=1.0):
        super().__init__()
        self.gain = gain
        self._engine = None

    def forward(self, lms: torch.Tensor) -> torch.Tensor:
        _check_cuda_f32(lms, "lms")
        x4, single = _as_batch(lms.contiguous())
        F, T = int(x4.shape[2]), int(x4.shape[3])
        if self._engine is None or self._engine.in_hw != (F, T):
            pl = ViewPlanner(mixup=False, rrc=False, rlf=True, n_global=1, in_hw=(F, T), canvas_hw=(F, T), fader_gain=self.gain)
            self._engine = ViewEngine(pl, (F, T), (F, T))
        plan = self._engine.planner.plan(x4.shape[0], device=x4.device)
        out = run_views(self._engine, x4, 0, F * T, plan, 1, 0, (F, T), None)[0]
        return out[0] if single else out

    def __repr__(self):
        return self.__class__.__name__ + f"(gain={self.gain})"


def log_mixup_exp(xa: torch.Tensor, xb: torch.Tensor, alpha: float) -> torch.Tensor:
    """reference: augmentations.py:81-85 -- log(alpha*exp(xa) + (1-alpha)*exp(xb) + eps), on the view kernel.
    xa, xb: (..., F, T) CUDA fp32 with F + T <= 256 and F*T a multiple of 4."""
    _check_cuda_f32(xa, "xa")
    _check_cuda_f32(xb, "xb")
    if xa.shape != xb.shape or xa.dim() < 2:
        raise ValueError("xa and xb must have the same shape (..., F, T)")
    F, T = int(xa.shape[-2]), int(xa.shape[-1])
    a2 = xa.contiguous().reshape(-1, F * T)
    b2 = xb.contiguous().reshape(-1, F * T)
    n = int(a2.shape[0])
    params = np.zeros((n, 1), dtype=VIEW_DTYPE)
    params["z_kind"] = 1
    params["z_index"][:, 0] = np.arange(n, dtype=np.int32)
    params["w_x"], params["w_z"] = np.float32(alpha), np.float32(1.0 - alpha)
    params["flags"] = 1
    out = torch.empty((n, 1, F, T), dtype=torch.float32, device=xa.device)
    if n:
        pdev = ViewEngine.upload(params, xa.device)
        _launch_views(_lib.load(), a2.data_ptr(), 0, F * T, b2, pdev.data_ptr(), 1, 0, n, 1, (F, T), (F, T), (F, T), [out], xa.device)
    return out.reshape(xa.shape)


class MixupBYOLA(_SingleStage):
    """Mixup for BYOL-A (reference: augmentations.py:88-122).

    `memory_bank` semantics are preserved: a FIFO of the last `n_memory` UN-mixed inputs; the mixing partner
    is `memory_bank[np.random.randint(len(memory_bank))]`.  The FIFO lives on the device as a ring.
    """

    def __init__(self, ratio=0.2, n_memory=2048, log_mixup_exp=True):
        super().__init__()
        if not log_mixup_exp:
            raise NotImplementedError("only log_mixup_exp=True is on the GPU path (the reference never uses False)")
        self.ratio = ratio
        self.n = n_memory
        self.log_mixup_exp = log_mixup_exp
        self._engine = None

    @property
    def memory_bank(self) -> List[torch.Tensor]:
        """The FIFO as a list of device tensors, oldest first (read-only view of the ring)."""
        eng = self._engine
        if eng is None or eng.ring is None:
            return []
        n = eng.planner.bank_len()
        last = self._pushed
        F, T = eng.in_hw
        return [eng.ring[(u % eng.planner.ring_slots)].view(1, F, T) for u in range(last - n, last)]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _check_cuda_f32(x, "x")
        x4, single = _as_batch(x.contiguous())
        F, T = int(x4.shape[2]), int(x4.shape[3])
        if self._engine is None:
            pl = ViewPlanner(mixup=True, rrc=False, rlf=False, mixup_ratio=self.ratio, n_memory=self.n,
                             ring_slots=self.n + max(1024, int(x4.shape[0])), n_global=1, in_hw=(F, T), canvas_hw=(F, T))
            self._engine = ViewEngine(pl, (F, T), (F, T))
            self._pushed = 0
        eng = self._engine
        eng.ensure_ring(x4.device)
        plan = eng.planner.plan(x4.shape[0], device=x4.device)
        out = run_views(eng, x4, 0, F * T, plan, 1, 0, (F, T), None)[0]
        push_bank(eng, x4, F * T, plan)
        self._pushed += int(x4.shape[0])
        return out[0] if single else out

    def __repr__(self):
        return self.__class__.__name__ + f"(ratio={self.ratio},n={self.n},log_mixup_exp={self.log_mixup_exp})"


class MixGaussianNoise(_SingleStage):
    """Gaussian Noise Mixer (reference: augmentations.py:125-140): `log((1 - lambd) * exp(lms) + exp(n * lambd) + eps)` with
    `lambd = ratio * np.random.rand()` replayed from numpy's global generator and n ~ N(0, 1).

    The reference takes n from torch's CPU generator (`torch.normal(0, lambd, shape)`); a GPU kernel cannot replay that stream, so
    `forward(lms, noise=None)` takes the standard-normal draws as an optional tensor of lms's shape -- pass the reference's draws
    (`torch.randn(shape)` after the same `torch.manual_seed`) to reproduce it exactly, or leave it out to draw from torch's CUDA
    generator: same distribution, different stream."""

    def __init__(self, ratio=0.2):
        super().__init__()
        self.ratio = ratio
        self._engine = None

    def forward(self, lms: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        _check_cuda_f32(lms, "lms")
        x4, single = _as_batch(lms.contiguous())
        F, T = int(x4.shape[2]), int(x4.shape[3])
        if self._engine is None or self._engine.in_hw != (F, T):
            pl = ViewPlanner(mixup=False, rrc=False, rlf=False, n_global=1, in_hw=(F, T), canvas_hw=(F, T), gnoise=True, gnoise_ratio=self.ratio)
            self._engine = ViewEngine(pl, (F, T), (F, T))
        if noise is not None:
            noise = noise.contiguous().reshape(x4.shape)
        plan = self._engine.planner.plan(x4.shape[0], device=x4.device)
        out = run_views(self._engine, x4, 0, F * T, plan, 1, 0, (F, T), None, noise)[0]
        return out[0] if single else out

    def __repr__(self):
        return self.__class__.__name__ + f"(ratio={self.ratio})"


class NormalizeBatch(nn.Module):
    """Normalization of an input batch (reference: augmentations.py:217-232, used per crop at main.py:62-66 under --post_norm):
    `(X - X.mean(axis)) / clamp(X.std(axis), eps)` with `axis=[0, 2, 3]` -- per-channel statistics over batch, frequency and time,
    unbiased std -- as two CUDA launches (abt_normalize_batch).  X: (B, C, F, T) CUDA fp32."""

    def __init__(self, axis=[0, 2, 3]):
        super().__init__()
        if list(axis) != [0, 2, 3]:
            raise NotImplementedError("only axis=[0, 2, 3] (the reference's only setting) is on the GPU path")
        self.axis = axis
        self._ws = {}

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        _check_cuda_f32(X, "X")
        if X.dim() != 4:
            raise ValueError(f"expected a batch (B, C, F, T), got {tuple(X.shape)}")
        X = X.contiguous()
        B, Cn, F, T = (int(v) for v in X.shape)
        out = torch.empty_like(X)
        lib = _lib.load()
        key = (X.device.index, Cn)
        if key not in self._ws:
            nbytes = C.c_size_t()
            _lib.check(lib.abt_normalize_batch_workspace_bytes(Cn, C.byref(nbytes)))
            self._ws[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(lib.abt_normalize_batch(X.data_ptr(), B, Cn, F * T, out.data_ptr(), self._ws[key].data_ptr(),
                                               torch.cuda.current_stream(X.device).cuda_stream))
        return out

    def __repr__(self):
        return self.__class__.__name__ + f"(axis={self.axis})"


class RunningNorm(nn.Module):
    """Online normalization using running mean / std over the samples (reference: augmentations.py:187-210, --pre_norm at
    main.py:272-277), default `axis=[1, 2]`: one scalar mean and std for the whole (1, F, T) log-mel, updated sample by sample up to
    `epoch_samples * max_update_epochs` samples.  A batch `(B, 1, F, T)` is processed in sample order, i.e. exactly as B consecutive
    calls of the reference module; the running state lives on the device (abt_running_norm).  x: CUDA fp32."""

    def __init__(self, epoch_samples, max_update_epochs=10, axis=[1, 2]):
        super().__init__()
        if list(axis) != [1, 2]:
            raise NotImplementedError("only axis=[1, 2] (the reference's only setting) is on the GPU path")
        self.max_update = epoch_samples * max_update_epochs
        self.axis = axis
        self._state = None
        self._ws = {}

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        _check_cuda_f32(image, "image")
        x4, single = _as_batch(image.contiguous())
        B, elems = int(x4.shape[0]), int(x4.shape[2] * x4.shape[3])
        dev = x4.device
        if self._state is None or self._state.device != dev:
            self._state = torch.zeros(3, dtype=torch.float64, device=dev)
        lib = _lib.load()
        if B not in self._ws:
            nbytes = C.c_size_t()
            _lib.check(lib.abt_running_norm_workspace_bytes(B, C.byref(nbytes)))
            self._ws[B] = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        out = torch.empty_like(x4)
        with torch.cuda.device(dev):
            _lib.check(lib.abt_running_norm(x4.data_ptr(), B, elems, int(self.max_update), self._state.data_ptr(), out.data_ptr(),
                                            self._ws[B].data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return out[0] if single else out

    def __repr__(self):
        return self.__class__.__name__ + f"(max_update={self.max_update},axis={self.axis})"
