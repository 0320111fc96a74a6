"""GPU audio frontend (reference: datasets.py, old/data_manager/wav_to_lms.py).

`LogMelSpectrogram` is the fused replacement of
    (torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, win_length, hop_length, n_mels,
                                          f_min, f_max, power=2)(wav) + torch.finfo().eps).log()
(datasets.py:39-48,115) with the optional dataset z-score (datasets.py:118-119) folded in.

`BatchFrontend` mirrors the arithmetic of `Dataset.__getitem__` for a whole batch that is already
on the GPU, in the reference's per-sample draw order:
  * lms path (AudioSet / --load_lms, datasets.py:336-357): log-mel of the whole clip, random
    96-frame crop drawn with np.random.randint, z-score, two views.  `mode="crop"` computes only the
    frames the crop needs (same result, 10x fewer FFTs for 10 s clips); `mode="full"` also materialises
    the (B, 64, T_full) log-mel, like the offline converter does.
  * wav path (--load_wav, datasets.py:98-122): centre pad to unit length, random unit crop drawn with
    random.randint, log-mel of that unit, z-score, two views.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .augmentations import ViewEngine
from .transforms import AudioPairTransform

__all__ = ["LogMelSpectrogram", "BatchFrontend"]


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class LogMelSpectrogram(nn.Module):
    """wav (..., L) CUDA fp32 -> log-mel (..., n_mels, 1 + L // hop_length) fp32."""

    def __init__(self, sample_rate=16000, n_fft=1024, win_length=None, hop_length=None, n_mels=64, f_min=0.0, f_max=None,
                 power=2, norm_stats: Optional[Sequence[float]] = None):
        super().__init__()
        if power != 2:
            raise ValueError("only power=2 (the reference's setting, datasets.py:47) is implemented")
        win_length = n_fft if win_length is None else win_length
        hop_length = win_length // 2 if hop_length is None else hop_length
        f_max = float(sample_rate // 2) if f_max is None else f_max
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = sample_rate, n_fft, win_length, hop_length
        self.n_mels, self.f_min, self.f_max = n_mels, f_min, f_max
        self.norm_stats = None if norm_stats is None else (float(norm_stats[0]), float(norm_stats[1]))
        self._plans = {}
        self._lib = _lib.load()

    def plan(self, device: torch.device):
        key = (device.type, device.index)
        if key not in self._plans:
            cfg = _lib.MelConfig()
            cfg.sample_rate, cfg.n_fft, cfg.win_length, cfg.hop_length = self.sample_rate, self.n_fft, self.win_length, self.hop_length
            cfg.n_mels, cfg.f_min, cfg.f_max = self.n_mels, float(self.f_min), float(self.f_max)
            cfg.apply_norm = int(self.norm_stats is not None)
            cfg.norm_mean, cfg.norm_std = self.norm_stats if self.norm_stats is not None else (0.0, 1.0)
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(self._lib.abt_logmel_plan_create(C.byref(cfg), C.byref(h)))
            self._plans[key] = h
        return self._plans[key]

    def __del__(self):
        for h in getattr(self, "_plans", {}).values():
            try:
                self._lib.abt_logmel_plan_destroy(h)
            except Exception:
                pass

    def n_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop_length

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        if not wav.is_cuda:
            raise RuntimeError("wav must be a CUDA tensor: ssl_audio_b200 has no CPU path")
        if wav.dtype != torch.float32:
            raise ValueError("wav must be float32")
        lead = wav.shape[:-1]
        L = int(wav.shape[-1])
        w2 = wav.contiguous().reshape(-1, L)
        B = int(w2.shape[0])
        T = self.n_frames(L)
        out = torch.empty((B, self.n_mels, T), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _lib.check(self._lib.abt_logmel_fwd(self.plan(wav.device), w2.data_ptr(), B, L, out.data_ptr(), _stream(wav.device)))
        return out.reshape(*lead, self.n_mels, T)

    def crop_into(self, wav2d: torch.Tensor, n_samples: int, wav_offset_ptr: int, frame_start_ptr: int, n_frames: int,
                  out_base: torch.Tensor, out_slot_ptr: int, out_slot_stride: int) -> None:
        """Crop-first log-mel of `n_frames` frames per clip written to out_base[out_slot[b]] (see abt_logmel_crop_fwd).
        The *_ptr arguments are device pointers to int32 arrays (0 = absent)."""
        B = int(wav2d.shape[0])
        with torch.cuda.device(wav2d.device):
            _lib.check(self._lib.abt_logmel_crop_fwd(
                self.plan(wav2d.device), wav2d.data_ptr(), int(wav2d.stride(0)), wav_offset_ptr or None, B, int(n_samples),
                frame_start_ptr or None, int(n_frames), out_base.data_ptr(), out_slot_ptr or None, int(out_slot_stride),
                _stream(wav2d.device)))


class BatchFrontend(nn.Module):
    """Batched `Dataset.__getitem__` arithmetic: waveforms (or precomputed log-mels) -> two augmented views.

    Args:
        cfg: the reference's `args` namespace (sample_rate, n_fft, win_length, hop_length, n_mels, f_min, f_max,
             unit_sec, crop_frames, + the AudioPairTransform fields).
        transform: an `AudioPairTransform` (built from cfg if None).
        norm_stats: dataset (mean, std), e.g. AudioSet (-0.8294, 4.6230) (main.py:293).
        path: "lms" (datasets.py:336-357 semantics) or "wav" (datasets.py:98-122 semantics).
        mode: "crop" (crop-first) or "full" (materialise the full log-mel; lms path only).
    """

    def __init__(self, cfg, transform: Optional[AudioPairTransform] = None, norm_stats: Optional[Sequence[float]] = None,
                 path: str = "lms", mode: str = "crop"):
        super().__init__()
        if path not in ("lms", "wav") or mode not in ("crop", "full"):
            raise ValueError("path must be 'lms' or 'wav', mode 'crop' or 'full'")
        self.cfg = cfg
        self.path, self.mode = path, mode
        self.transform = transform if transform is not None else AudioPairTransform(cfg)
        self.norm_stats = None if norm_stats is None else (float(norm_stats[0]), float(norm_stats[1]))
        self.crop_frames = int(cfg.crop_frames)
        self.unit_length = int(cfg.unit_sec * cfg.sample_rate)
        mel_kw = dict(sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, win_length=cfg.win_length, hop_length=cfg.hop_length,
                      n_mels=cfg.n_mels, f_min=cfg.f_min, f_max=cfg.f_max, power=2)
        self.logmel_norm = LogMelSpectrogram(norm_stats=self.norm_stats, **mel_kw)   # crop-first: z-score fused
        self.logmel_raw = LogMelSpectrogram(norm_stats=None, **mel_kw)               # full mode: raw log-mel, as stored in .npy
        self.last_lms: Optional[torch.Tensor] = None
        self._lib = _lib.load()

    # -- precomputed log-mel input (the reference's default --load_lms path) --------------------
    def forward_lms(self, lms: torch.Tensor) -> List[torch.Tensor]:
        """lms (B, n_mels, T_full) CUDA fp32, un-normalised -> list of views."""
        if not lms.is_cuda or lms.dtype != torch.float32 or lms.dim() != 3:
            raise ValueError("lms must be a CUDA float32 tensor (B, n_mels, T_full)")
        lms = lms.contiguous()
        B, F, T_full = (int(v) for v in lms.shape)
        tf = self.transform
        eng = tf.engine(B)
        ring = eng.ensure_ring(lms.device)
        crop_range = T_full - self.crop_frames if T_full > self.crop_frames else 0
        plan = eng.planner.plan(B, time_crop_range=crop_range, device=lms.device)
        mean, std = self.norm_stats if self.norm_stats is not None else (0.0, 1.0)
        with torch.cuda.device(lms.device):
            _lib.check(self._lib.abt_lms_crop_norm(lms.data_ptr(), B, F, T_full, plan.starts_ptr, self.crop_frames,
                                                   int(self.norm_stats is not None), mean, std, ring.data_ptr(), plan.slots_ptr,
                                                   int(ring.shape[1]), _stream(lms.device)))
        self.last_plan = plan
        return tf.views_from_plan(ring, plan.slots_ptr, int(ring.shape[1]), plan)

    # -- two-step form of the crop-first path: host work (and the PCIe transfer) first, launches later ---------------
    def prepare(self, wav: torch.Tensor, device: Optional[torch.device] = None, static: bool = False):
        """Host half of `forward` (path "lms", mode "crop"): draws the batch's random parameters in the reference's order and
        uploads them (one async copy on the current stream).  `launch(handle)` then only enqueues the two kernels, so a trainer
        can plan early and launch exactly where it wants the kernels to overlap something else.

        With waveforms in PINNED host memory, `prepare` also enqueues the crop-first span gather (the only PCIe traffic of the
        batch) on the current stream, into one of two alternating span buffers: calling `prepare(next_batch)` on a side stream
        while the current batch computes is the prefetch a `DataLoader` worker would otherwise provide.  `launch` makes its stream
        wait for that gather.

        `static=True` (device waveforms): the plan is uploaded into ONE fixed device buffer, so that `launch(handle)` can be captured
        in a CUDA graph once and replayed after every later `prepare(wav, static=True)` on the same stream (same `wav` storage)."""
        if wav.dtype != torch.float32 or wav.dim() != 2:
            raise ValueError("wav must be a float32 tensor (B, L)")
        if self.path != "lms" or self.mode != "crop":
            raise ValueError("prepare/launch cover the crop-first lms path only")
        if not wav.is_cuda:
            if static:
                raise ValueError("static plans are for device waveforms (graph capture); host waveforms use the prefetch form")
            return self._prepare_host(wav, device)
        wav = wav.contiguous()
        B, L = int(wav.shape[0]), int(wav.shape[1])
        eng = self.transform.engine(B)
        ring = eng.ensure_ring(wav.device)
        T_full = self.logmel_raw.n_frames(L)
        crop_range = T_full - self.crop_frames if T_full > self.crop_frames else 0
        plan = eng.planner.plan(B, time_crop_range=crop_range, device=wav.device, static=static)
        return (wav, L, ring, plan)

    def _prepare_host(self, wav: torch.Tensor, device: Optional[torch.device]):
        if not wav.is_contiguous():
            raise ValueError("wav must be a contiguous float32 host tensor (B, L)")
        if not wav.is_pinned():
            raise RuntimeError("host waveforms must be in pinned memory (DataLoader(pin_memory=True) or tensor.pin_memory()); "
                               "ssl_audio_b200 has no pageable-memory / CPU path")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B, L = int(wav.shape[0]), int(wav.shape[1])
        lib = self._lib
        span_len = C.c_int()
        _lib.check(lib.abt_wav_span_len(self.logmel_norm.plan(dev), self.crop_frames, C.byref(span_len)))
        T_full = self.logmel_raw.n_frames(L)
        if T_full <= self.crop_frames or L < span_len.value or B == 0:      # nothing to skip: the clips are copied whole
            with torch.cuda.device(dev):
                return self.prepare(wav.to(dev, non_blocking=True))
        eng = self.transform.engine(B)
        ring = eng.ensure_ring(dev)
        key = (dev.index, B, span_len.value)
        if getattr(self, "_span_key", None) != key:
            self._span_bufs = [(torch.empty((B, span_len.value), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
                                torch.cuda.Event(), torch.cuda.Event()) for _ in range(2)]      # spans, origins, gathered, consumed
            self._span_next = 0
            self._span_key = key
        k = self._span_next
        self._span_next = k ^ 1
        spans, origin, gathered, consumed = self._span_bufs[k]
        with torch.cuda.device(dev):
            st_t = torch.cuda.current_stream(dev)
            st_t.wait_event(consumed)            # the log-mel launch that last read this buffer (two batches ago)
            plan = eng.planner.plan(B, time_crop_range=T_full - self.crop_frames, device=dev)
            _lib.check(lib.abt_wav_span_gather(self.logmel_norm.plan(dev), wav.data_ptr(), 1, L, B, L, plan.starts_ptr, self.crop_frames,
                                               spans.data_ptr(), origin.data_ptr(), _stream(dev)))
            gathered.record(st_t)
        self.h2d_bytes = B * span_len.value * 4
        return ("host", wav, L, ring, plan, k, dev)

    def launch(self, handle) -> List[torch.Tensor]:
        host = isinstance(handle[0], str)                  # ("host", wav, L, ring, plan, k, dev) or (wav, L, ring, plan)
        handle[4 if host else 3].check_live()              # the plan's staging slot must not have been reused
        if host:
            _, wav, L, ring, plan, k, dev = handle          # wav is kept alive until its gather has been consumed
            spans, origin, gathered, consumed = self._span_bufs[k]
            stride = int(ring.shape[1])
            B = int(spans.shape[0])
            with torch.cuda.device(dev):
                st_t = torch.cuda.current_stream(dev)
                st_t.wait_event(gathered)
                _lib.check(self._lib.abt_logmel_span_fwd(self.logmel_norm.plan(dev), spans.data_ptr(), origin.data_ptr(), B, L, plan.starts_ptr,
                                                         self.crop_frames, ring.data_ptr(), plan.slots_ptr, stride, _stream(dev)))
                consumed.record(st_t)
                self.last_plan = plan
                return self.transform.views_from_plan(ring, plan.slots_ptr, stride, plan)
        wav, L, ring, plan = handle
        stride = int(ring.shape[1])
        self.logmel_norm.crop_into(wav, L, 0, plan.starts_ptr, self.crop_frames, ring, plan.slots_ptr, stride)
        self.last_plan = plan
        return self.transform.views_from_plan(ring, plan.slots_ptr, stride, plan)

    # -- waveform input ---------------------------------------------------------------------------
    def forward(self, wav: torch.Tensor) -> List[torch.Tensor]:
        """wav (B, L) fp32, CUDA or PINNED host memory -> list of views [(B,1,F,T), (B,1,F,T), local crops...]."""
        if wav.dtype != torch.float32 or wav.dim() != 2:
            raise ValueError("wav must be a float32 tensor (B, L)")
        if not wav.is_cuda:
            return self.forward_host(wav)
        wav = wav.contiguous()
        B, L = int(wav.shape[0]), int(wav.shape[1])
        tf = self.transform
        eng = tf.engine(B)
        ring = eng.ensure_ring(wav.device)
        stride = int(ring.shape[1])
        if self.path == "lms":
            T_full = self.logmel_raw.n_frames(L)
            if self.mode == "full":
                self.last_lms = self.logmel_raw(wav)
                return self.forward_lms(self.last_lms)
            return self.launch(self.prepare(wav))
        else:
            # datasets.py:103-113: centre pad to unit_length, then random.randint unit crop
            if L < self.unit_length:
                adj = self.unit_length - L
                wav = torch.nn.functional.pad(wav, (adj // 2, adj - adj // 2))
                L = self.unit_length
            plan = eng.planner.plan(B, wav_crop_range=L - self.unit_length, device=wav.device)
            n_frames = self.logmel_norm.n_frames(self.unit_length)
            if n_frames != self.crop_frames:
                raise ValueError(f"unit_sec gives {n_frames} frames but crop_frames is {self.crop_frames}")
            self.logmel_norm.crop_into(wav, self.unit_length, plan.wav_starts_ptr, 0, n_frames, ring, plan.slots_ptr, stride)
        self.last_plan = plan
        return tf.views_from_plan(ring, plan.slots_ptr, stride, plan)

    # -- waveforms in pinned host memory (what a DataLoader with pin_memory=True hands over, main.py:308-309) ---------
    def forward_host(self, wav: torch.Tensor, device: Optional[torch.device] = None) -> List[torch.Tensor]:
        """wav (B, L) fp32 in PINNED host memory -> views on `device` (default: the current CUDA device).

        Crop-first all the way to the host: the crop is planned first, then a small kernel reads ONLY the samples the
        cropped frames need straight out of the pinned buffer over PCIe (abt_wav_span_gather), so a 10 s clip costs
        65 KB of host->device traffic instead of 640 KB.  Results are bit-identical to `forward(wav.cuda())`.
        Clips that are not longer than the crop (nothing to skip) are copied whole.  `prepare(wav)` / `launch(handle)` is the
        same thing in two steps (prefetch the next batch's samples while this one computes)."""
        if wav.is_cuda or wav.dtype != torch.float32 or wav.dim() != 2 or not wav.is_contiguous():
            raise ValueError("wav must be a contiguous float32 host tensor (B, L)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.path != "lms" or self.mode != "crop":
            if not wav.is_pinned():
                raise RuntimeError("host waveforms must be in pinned memory (DataLoader(pin_memory=True) or tensor.pin_memory()); "
                                   "ssl_audio_b200 has no pageable-memory / CPU path")
            return self.forward(wav.to(dev, non_blocking=True))
        return self.launch(self.prepare(wav, dev))
