"""Host planner wrapper: draws the reference's random parameters for a whole batch.

The reference draws per sample, inside `Dataset.__getitem__`, from numpy's global legacy
generator and CPython's `random` (augmentations.py:34-37,70,105,108; datasets.py:89,112,344).
`ViewPlanner.plan` imports the state of those two global generators into the native planner
(csrc/planner.cpp), lets it replay the exact draw order for `n_clips` samples, and writes the
advanced states back, so `np.random.seed(s); random.seed(s)` means what it means for the
reference run with `num_workers=0`.
"""
from __future__ import annotations

import ctypes as C
import random as _pyrandom
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib

VIEW_DTYPE = np.dtype([
    ("z_kind", np.int32), ("z_index", np.int32), ("w_x", np.float32), ("w_z", np.float32),
    ("i", np.int32), ("j", np.int32), ("h", np.int32), ("w", np.int32),
    ("head", np.float32), ("tail", np.float32), ("flags", np.int32), ("out_index", np.int32),
    ("g_lambda", np.float32), ("g_keep", np.float32), ("reserved", np.int32, (2,)),
])
assert VIEW_DTYPE.itemsize == C.sizeof(_lib.ViewParams) == 64

FLAG_MIXUP, FLAG_RRC, FLAG_RLF, FLAG_GNOISE = 1, 2, 4, 8


def _numpy_global_state_address() -> int:
    bitgen = np.random.mtrand._rand._bit_generator
    if type(bitgen).__name__ != "MT19937":
        raise RuntimeError("numpy's global RandomState is not MT19937-backed; cannot replay the reference's draws")
    return int(bitgen.ctypes.state_address)


_PYRANDOM_ADDR = None   # (address of `int index`, address of `uint32_t state[624]`) inside random._inst, or False


def _pyrandom_state_address():
    """CPython keeps the Mersenne Twister of the `random` module inside the `random._inst` object as
    `struct { PyObject_HEAD; int index; uint32_t state[624]; }` (Modules/_randommodule.c).  Reading and writing it in place
    saves two 625-element tuple conversions per batch.  The layout is VERIFIED against random.getstate() before it is
    trusted; on any mismatch (other interpreter, other layout) the slow getstate()/setstate() path is used."""
    global _PYRANDOM_ADDR
    if _PYRANDOM_ADDR is None:
        _PYRANDOM_ADDR = False
        try:
            inst = _pyrandom._inst
            base = id(inst) + C.sizeof(C.c_ssize_t) + C.sizeof(C.c_void_p)
            st = inst.getstate()[1]
            idx = C.c_int.from_address(base).value
            key = (C.c_uint32 * 624).from_address(base + 4)
            if idx == st[624] and tuple(key) == tuple(st[:624]):
                _PYRANDOM_ADDR = (base, base + 4)
        except Exception:
            _PYRANDOM_ADDR = False
    return _PYRANDOM_ADDR


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_LOCK = _NoLock()


@dataclass
class BatchPlan:
    starts: np.ndarray       # (B,) int32 time-crop start per clip, -1 if none was drawn
    wav_starts: np.ndarray   # (B,) int32 wav-crop start per clip, -1 if none was drawn
    params: np.ndarray       # (B, n_views) VIEW_DTYPE
    slots: np.ndarray        # (B,) int32 bank ring slot of each clip
    # device copy of the packed plan (one H2D copy), when planned with a device
    dev: Optional["torch.Tensor"] = None
    params_ptr: int = 0
    starts_ptr: int = 0
    wav_starts_ptr: int = 0
    slots_ptr: int = 0
    # staging-ring bookkeeping: the host arrays above and the device copy alias ring slot `_gen % depth`, which is handed out
    # again `depth` plans later -- `check_live()` refuses a plan whose slot has been reused
    _staging: Optional["PlanStaging"] = None
    _gen: int = -1

    def check_live(self):
        st = self._staging
        if st is not None and st.gen - self._gen >= len(st.host):
            raise RuntimeError(f"stale batch plan: {st.gen - self._gen} plans were drawn since this one and its staging slot "
                               f"(ring depth {len(st.host)}) has been reused; launch a prepared batch before preparing {len(st.host)} more")


class PlanStaging:
    """Ring of pinned host buffers + device buffers for the packed plan: one cudaMemcpyAsync per batch and no
    per-step pinned allocation.  A slot is reused only after the copy that read it has completed."""

    def __init__(self, device, nbytes: int, depth: int = 4):
        import torch
        self.device = device
        self.nbytes = int(nbytes)
        self.host = [torch.empty(self.nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_np = [h.numpy() for h in self.host]
        self.dev = [torch.empty(self.nbytes, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.done = [None] * depth
        self.k = 0
        self.gen = 0            # plans handed out so far
        self.dev_static = None

    def acquire(self):
        k = self.k
        self.k = (k + 1) % len(self.host)
        self.gen += 1
        if self.done[k] is not None:
            self.done[k].synchronize()
        return k

    def upload(self, k: int, nbytes: int, static: bool = False):
        """`static`: always the same device buffer (for kernels captured in a CUDA graph, whose pointers are frozen): the copy is ordered
        on the current stream, so a graph replayed on that stream afterwards reads this plan and the next upload waits for the replay."""
        import torch
        if static and self.dev_static is None:
            self.dev_static = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        dst = self.dev_static if static else self.dev[k]
        dst[:nbytes].copy_(self.host[k][:nbytes], non_blocking=True)
        ev = self.done[k] if self.done[k] is not None else torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.done[k] = ev
        return dst


class ViewPlanner:
    def __init__(self, *, mixup: bool, rrc: bool, rlf: bool, mixup_ratio: float = 0.2, n_memory: int = 2048,
                 ring_slots: Optional[int] = None, n_global: int = 2, in_hw: Tuple[int, int] = (64, 96),
                 canvas_hw: Tuple[int, int] = (64, 144), freq_scale: Sequence[float] = (0.6, 1.5),
                 time_scale: Sequence[float] = (0.6, 1.5), n_local: int = 0, local_hw: Tuple[int, int] = (16, 16),
                 local_scale: Sequence[float] = (0.05, 0.6), fader_gain: float = 1.0, gnoise: bool = False,
                 gnoise_ratio: float = 0.2):
        lib = _lib.load()
        cfg = _lib.PlanConfig()
        cfg.mixup, cfg.rrc, cfg.rlf = int(bool(mixup)), int(bool(rrc)), int(bool(rlf))
        cfg.mixup_ratio_d = float(mixup_ratio)
        cfg.n_memory = int(n_memory)
        cfg.ring_slots = int(ring_slots if ring_slots is not None else n_memory + 1024)
        cfg.n_global = int(n_global)
        cfg.in_h, cfg.in_w = int(in_hw[0]), int(in_hw[1])
        cfg.canvas_h, cfg.canvas_w = int(canvas_hw[0]), int(canvas_hw[1])
        cfg.freq_scale[0], cfg.freq_scale[1] = float(freq_scale[0]), float(freq_scale[1])
        cfg.time_scale[0], cfg.time_scale[1] = float(time_scale[0]), float(time_scale[1])
        cfg.n_local = int(n_local)
        cfg.local_h, cfg.local_w = int(local_hw[0]), int(local_hw[1])
        cfg.local_scale[0], cfg.local_scale[1] = float(local_scale[0]), float(local_scale[1])
        cfg.fader_gain = float(fader_gain)
        cfg.gnoise = int(bool(gnoise))
        cfg.gnoise_ratio_d = float(gnoise_ratio)
        self.cfg = cfg
        self.n_views = cfg.n_global + cfg.n_local
        self.ring_slots = cfg.ring_slots
        self.n_memory = cfg.n_memory
        self._uses_numpy = bool(mixup or rrc or rlf or n_local or gnoise)
        self.gnoise = bool(gnoise)
        self._uses_pyrandom = bool(rrc or n_local)
        self._h = C.c_void_p()
        _lib.check(lib.abt_planner_create(C.byref(cfg), C.byref(self._h)))
        self._lib = lib
        self._key = np.empty(624, dtype=np.uint32)
        self._staging = {}
        self._layouts = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.abt_planner_destroy(h)
            self._h = C.c_void_p()

    # -- bank bookkeeping (mirrors len(MixupBYOLA.memory_bank)) -------------------------------
    def bank_len(self) -> int:
        return int(self._lib.abt_planner_bank_len(self._h))

    def bank_reset(self) -> None:
        self._lib.abt_planner_bank_reset(self._h)

    # -- planning ------------------------------------------------------------------------------
    def plan(self, n_clips: int, time_crop_range: int = 0, wav_crop_range: int = 0, device=None, static: bool = False) -> BatchPlan:
        """Draw the parameters of `n_clips` samples.  With `device`, the packed plan is also uploaded to that CUDA
        device (one asynchronous copy on the current stream) and the device pointers are filled in.  `static`: the upload goes to ONE
        fixed device buffer (CUDA-graph capture of the consuming kernels); such a plan is valid until the next static plan."""
        lib, h = self._lib, self._h
        use_np = self._uses_numpy or time_crop_range > 0
        use_py = self._uses_pyrandom or wav_crop_range > 0
        lay = self._layouts.get(n_clips)
        if lay is None:
            total, o1, o2, o3 = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
            _lib.check(lib.abt_planner_packed_bytes(h, n_clips, C.byref(total), C.byref(o1), C.byref(o2), C.byref(o3)))
            lay = self._layouts[n_clips] = (total.value, o1.value, o2.value, o3.value)
        nbytes, off1, off2, off3 = lay
        slot = -1
        if device is not None:
            st = self._staging.get(device)
            if st is None or st.nbytes < nbytes:
                st = PlanStaging(device, max(nbytes, 1 << 16))
                self._staging[device] = st
            slot = st.acquire()
            buf = st.host_np[slot]
        else:
            buf = np.empty(max(nbytes, 16), dtype=np.uint8)
        # the native planner advances numpy's global MT19937 state in place: hold numpy's own lock so that another thread drawing
        # from np.random at the same time cannot interleave (CPython's `random` has no lock: keep its use single-threaded)
        np_lock = np.random.mtrand._rand._bit_generator.lock if use_np else _NO_LOCK
        np_addr = _numpy_global_state_address() if use_np else None
        py_addr = _pyrandom_state_address() if use_py else None
        if use_py and not py_addr:
            # unknown interpreter layout: go through random.getstate() / setstate()
            pst = _pyrandom.getstate()
            pkey = np.array(pst[1][:624], dtype=np.uint32)
            pidx = C.c_int32(int(pst[1][624]))
            with np_lock:
                _lib.check(lib.abt_planner_plan_batch_global(h, n_clips, int(time_crop_range), int(wav_crop_range), np_addr, C.addressof(pidx),
                                                             pkey.ctypes.data, buf.ctypes.data, buf.nbytes))
            _pyrandom.setstate((pst[0], tuple(pkey.tolist()) + (int(pidx.value),), pst[2]))
        else:
            with np_lock:
                _lib.check(lib.abt_planner_plan_batch_global(h, n_clips, int(time_crop_range), int(wav_crop_range), np_addr,
                                                             py_addr[0] if py_addr else None, py_addr[1] if py_addr else None,
                                                             buf.ctypes.data, buf.nbytes))
        nv = self.n_views
        params = buf[:VIEW_DTYPE.itemsize * nv * n_clips].view(VIEW_DTYPE).reshape(n_clips, nv)
        starts = buf[off1:off1 + 4 * n_clips].view(np.int32)
        wav_starts = buf[off2:off2 + 4 * n_clips].view(np.int32)
        slots = buf[off3:off3 + 4 * n_clips].view(np.int32)
        plan = BatchPlan(starts, wav_starts, params, slots)
        if device is not None:
            dev = st.upload(slot, nbytes, static)
            base = dev.data_ptr()
            plan.dev = dev
            if not static:
                plan._staging, plan._gen = st, st.gen
            plan.params_ptr, plan.starts_ptr, plan.wav_starts_ptr, plan.slots_ptr = base, base + off1, base + off2, base + off3
        return plan
