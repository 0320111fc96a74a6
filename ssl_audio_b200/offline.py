"""Offline / evaluation-time consumers of the log-mel kernel.

Mirrors, on the GPU path:
  * `ToLogMelSpec` + the `.npy` cache writer of old/data_manager/wav_to_lms.py:30-87 (float32 `(64, T_full)` arrays, one file per
    clip, existing files are left alone) -- `LogMelCacheWriter`;
  * `calculate_norm_stats` of datasets.py:362-376 (mean and std + eps over `n_norm_calc` randomly drawn samples) -- one reduction
    kernel (`abt_mean_std`) instead of stacking 10 000 tensors on the host;
  * the evaluation features of main.py:240-252 (`crop_frames=711`, `transform=None`): log-mel -> random 711-frame crop or right
    zero-pad (datasets.py:87-96) -> dataset z-score -- `EvalFeatures`;
  * the HEAR wrapper's input pipeline, hear/sample/vit.py:90-106 (`_to_feature`, `_normalize_batch`; `win_length: 400` in
    hear/config.yaml) -- `HearFeatures`.
File decoding / resampling (librosa) is I/O and stays with the caller: these classes take decoded waveforms.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .frontend import LogMelSpectrogram

__all__ = ["FFT_parameters", "ToLogMelSpec", "LogMelCacheWriter", "mean_std", "calculate_norm_stats", "EvalFeatures", "HearFeatures"]

F32_EPS = float(torch.finfo(torch.float32).eps)


class FFT_parameters:
    """old/data_manager/wav_to_lms.py:30-38."""
    sample_rate = 16000
    window_size = 1024
    n_fft = 1024
    hop_size = 160
    n_mels = 64
    f_min = 60
    f_max = 7800


def _as_cuda_wave(audio, device) -> torch.Tensor:
    t = torch.as_tensor(audio, dtype=torch.float32)
    return t.to(device, non_blocking=True).contiguous()


class ToLogMelSpec:
    """old/data_manager/wav_to_lms.py:41-61: `audio` (L,) or (B, L) -> log-mel (64, T) or (B, 64, T), a CUDA tensor."""

    def __init__(self, cfg=FFT_parameters, device: Optional[torch.device] = None):
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.to_spec = LogMelSpectrogram(sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, win_length=cfg.window_size, hop_length=cfg.hop_size,
                                         n_mels=cfg.n_mels, f_min=cfg.f_min, f_max=cfg.f_max, power=2)

    def __call__(self, audio) -> torch.Tensor:
        return self.to_spec(_as_cuda_wave(audio, self.device))


class LogMelCacheWriter:
    """The offline converter's per-file contract (wav_to_lms.py:64-87) for already decoded clips: `<to_dir>/<name minus suffix>.npy`
    holds the float32 (n_mels, T_full) log-mel; files that exist are skipped.  Clips of equal length are converted in one launch."""

    def __init__(self, to_dir: str, cfg=FFT_parameters, suffix: str = ".wav", device: Optional[torch.device] = None):
        self.to_dir, self.suffix = str(to_dir), suffix
        self.to_lms = ToLogMelSpec(cfg, device)

    def target(self, subpathname: str) -> str:
        stem = subpathname[:-len(self.suffix)] if self.suffix and subpathname.endswith(self.suffix) else subpathname
        return os.path.join(self.to_dir, stem + ".npy")

    def convert(self, subpathnames: Sequence[str], waves: Sequence) -> List[str]:
        """Returns the base names written ('' for skipped files), in input order -- what `_converter_worker` returns."""
        out = [""] * len(subpathnames)
        todo = [k for k, nme in enumerate(subpathnames) if not os.path.exists(self.target(nme))]
        by_len = {}
        for k in todo:
            by_len.setdefault(int(np.shape(waves[k])[-1]), []).append(k)
        for _, ks in sorted(by_len.items()):
            batch = torch.stack([torch.as_tensor(waves[k], dtype=torch.float32).reshape(-1) for k in ks])
            lms = self.to_lms(batch).cpu().numpy()
            for row, k in enumerate(ks):
                path = self.target(subpathnames[k])
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                np.save(path, lms[row])
                out[k] = os.path.basename(path)
        return out


def mean_std(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x.mean(), x.std()) of a CUDA float32 tensor as 0-dim float64 CUDA tensors (unbiased std, like torch.std)."""
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor: ssl_audio_b200 has no CPU path")
    if x.dtype != torch.float32:
        raise ValueError("x must be float32")
    x = x.contiguous()
    lib = _lib.load()
    n0 = int(x.shape[0]) if x.dim() > 0 else 1
    elems = x.numel() // max(n0, 1)
    if x.numel() == 0 or elems >= 2 ** 31:
        raise ValueError("empty tensor or more than 2^31 elements per row")
    with torch.cuda.device(x.device):
        nbytes = C.c_size_t()
        _lib.check(lib.abt_normalize_batch_workspace_bytes(1, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        out = torch.empty(2, dtype=torch.float64, device=x.device)
        _lib.check(lib.abt_mean_std(x.data_ptr(), n0, elems, out.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    return out[0], out[1]


def calculate_norm_stats(dataset, n_norm_calc: int = 10000, json_path: Optional[str] = "norm_stats.json", device: Optional[torch.device] = None,
                         chunk: int = 1024):
    """datasets.py:362-376: draw `n_norm_calc` indices with np.random.randint (same draw), take `dataset[i][0]`, return
    (mean, std + eps) over all of them.  The samples are reduced on the GPU chunk by chunk (exact combination of the chunks' sums
    in float64), not stacked on the host."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idxs = np.random.randint(0, len(dataset), size=n_norm_calc)
    n_tot, s1, s2 = 0.0, 0.0, 0.0
    for c0 in range(0, len(idxs), chunk):
        part = torch.stack([torch.as_tensor(dataset[int(i)][0], dtype=torch.float32) for i in idxs[c0:c0 + chunk]]).to(dev)
        m, sd = mean_std(part)
        n = float(part.numel())
        m, sd = float(m), float(sd)
        # chunk -> running totals (sum, sum of squared deviations), combined exactly (Chan et al.)
        css = sd * sd * (n - 1.0)
        if n_tot == 0.0:
            n_tot, s1, s2 = n, m * n, css
        else:
            delta = m - s1 / n_tot
            s2 += css + delta * delta * n_tot * n / (n_tot + n)
            s1 += m * n
            n_tot += n
    mean = s1 / n_tot
    std = float(np.sqrt(s2 / (n_tot - 1.0))) if n_tot > 1 else 0.0
    norm_stats = (float(np.float32(mean)), float(np.float32(std)) + F32_EPS)
    if json_path:
        with open(json_path, mode="w") as jsonfile:
            json.dump({"mean": norm_stats[0], "std": norm_stats[1]}, jsonfile, indent=2)
    return norm_stats


class EvalFeatures:
    """Evaluation-time inputs (main.py:240-252: `datasets.FSD50K(..., transform=None, norm_stats=..., crop_frames=711)`): per clip
    log-mel -> random `crop_frames` crop (np.random.randint, datasets.py:89) or right zero-pad (datasets.py:93-95) -> z-score.
    Takes waveforms (B, L) or precomputed log-mels (B, n_mels, T_full); returns (B, 1, n_mels, crop_frames) on the GPU."""

    def __init__(self, cfg, norm_stats: Optional[Sequence[float]] = None, crop_frames: int = 711):
        self.cfg, self.crop_frames = cfg, int(crop_frames)
        self.norm_stats = None if norm_stats is None else (float(norm_stats[0]), float(norm_stats[1]))
        self.logmel = LogMelSpectrogram(sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, win_length=cfg.win_length, hop_length=cfg.hop_length,
                                        n_mels=cfg.n_mels, f_min=cfg.f_min, f_max=cfg.f_max, power=2, norm_stats=self.norm_stats)
        self._lib = _lib.load()

    def _starts(self, b: int, t_full: int, device) -> Optional[torch.Tensor]:
        if t_full <= self.crop_frames:
            return None
        starts = np.array([np.random.randint(t_full - self.crop_frames) for _ in range(b)], dtype=np.int32)     # one draw per clip, in order
        return torch.from_numpy(starts).to(device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("inputs must be CUDA float32 tensors: ssl_audio_b200 has no CPU path")
        x = x.contiguous()
        dev = x.device
        b = int(x.shape[0])
        n_mels = int(self.cfg.n_mels)
        out = torch.empty((b, 1, n_mels, self.crop_frames), dtype=torch.float32, device=dev)
        stride = n_mels * self.crop_frames
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            if x.dim() == 2:            # waveforms: crop-first (only the needed frames are computed; padding frames get the z-scored zero)
                t_full = self.logmel.n_frames(int(x.shape[1]))
                starts = self._starts(b, t_full, dev)
                self.logmel.crop_into(x, int(x.shape[1]), 0, starts.data_ptr() if starts is not None else 0, self.crop_frames, out, 0, stride)
            elif x.dim() == 3:
                t_full = int(x.shape[2])
                starts = self._starts(b, t_full, dev)
                mean, std = self.norm_stats if self.norm_stats is not None else (0.0, 1.0)
                _lib.check(self._lib.abt_lms_crop_norm(x.data_ptr(), b, n_mels, t_full, starts.data_ptr() if starts is not None else None, self.crop_frames,
                                                       int(self.norm_stats is not None), mean, std, out.data_ptr(), None, stride, st))
            else:
                raise ValueError("expected waveforms (B, L) or log-mels (B, n_mels, T)")
        return out


class HearFeatures:
    """hear/sample/vit.py:90-106: `_to_feature` (log-mel of the batch, unsqueeze(1)) and `_normalize_batch` ((x - x.mean()) / x.std(),
    one mean / unbiased std over the whole batch)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.to_melspec = LogMelSpectrogram(sample_rate=cfg.sample_rate, n_fft=cfg.n_fft, win_length=cfg.win_length, hop_length=cfg.hop_length,
                                            n_mels=cfg.n_mels, f_min=cfg.f_min, f_max=cfg.f_max, power=2)
        self._lib = _lib.load()

    def _to_feature(self, batch_audio: torch.Tensor) -> torch.Tensor:
        return self.to_melspec(batch_audio).unsqueeze(1)

    def _normalize_batch(self, x: torch.Tensor) -> torch.Tensor:
        from .augmentations import NormalizeBatch
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError("expected (B, 1, n_mels, T)")
        return NormalizeBatch()(x)                 # one channel: per-channel statistics over (batch, freq, time) = the global ones

    def _to_normalized_spec(self, batch_audio: torch.Tensor) -> torch.Tensor:
        return self._normalize_batch(self._to_feature(batch_audio))

    def _get_timestamps(self, batch_audio: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        audio_len = int(batch_audio.shape[1])
        sec = audio_len / self.cfg.sample_rate
        x_len = int(x.shape[1])
        step = sec / x_len
        ts = torch.tensor([step * i for i in range(x_len)]).unsqueeze(0)
        return ts.repeat(int(batch_audio.shape[0]), 1)
