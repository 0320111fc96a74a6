"""CPU baseline port -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

The reference (jonahanton/SSL_audio) cannot travel to the GPU box (/root/reference is absent there and
it is Python, so there is nothing to compile into oracle/_ref).  This module restates its hot path with
the SAME library calls the reference makes on the CPU (torchaudio MelSpectrogram, torch exp/log,
F.interpolate(bicubic, align_corners=True), torch.linspace, nn.BatchNorm1d + matmul + autograd), so that
`bench.py --impl reference` and the `cpu_baseline` leg time what the reference's DataLoader workers and
loss module would execute.  It is pinned against the golden fixtures in tests/test_torch_port.py.

Only bench.py (cpu_baseline / --impl reference) and tests/ may import this file.
"""
from __future__ import annotations

import random
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = torch.finfo(torch.float32).eps


def make_melspec(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160, n_mels=64, f_min=60.0, f_max=7800.0):
    """The transform the reference builds at datasets.py:39-48."""
    import torchaudio.transforms as AT
    return AT.MelSpectrogram(sample_rate=sample_rate, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                             n_mels=n_mels, f_min=f_min, f_max=f_max, power=2)


def log_mel(melspec, wav: torch.Tensor) -> torch.Tensor:
    """datasets.py:115 / old/data_manager/wav_to_lms.py:58-61."""
    return (melspec(wav) + EPS).log()


class PortPairTransform:
    """Per-sample two-view transform in the reference's op order (utils/transforms.py:49-58 driving
    augmentations.py:40-55, 69-74, 81-85, 103-117) with default flags (mixup, RRC, RLF on)."""

    def __init__(self, out_size=(64, 96), vcs=(1.0, 1.5), scale=(0.6, 1.5), ratio=0.2, n_memory=2048):
        self.out_size, self.vcs, self.scale, self.ratio, self.n = out_size, vcs, scale, ratio, n_memory
        self.bank: List[torch.Tensor] = []

    def _mix(self, x):
        alpha = self.ratio * np.random.random()
        if self.bank:
            z = self.bank[np.random.randint(len(self.bank))]
            a = 1.0 - alpha
            mixed = torch.log(a * x.exp() + (1.0 - a) * z.exp() + EPS)
        else:
            mixed = x
        self.bank = (self.bank + [x])[-self.n:]
        return mixed.to(torch.float)

    def _rrc(self, x):
        _, fh, tw = x.shape
        ch, cw = int(fh * self.vcs[0]), int(tw * self.vcs[1])
        canvas = torch.zeros((1, ch, cw), dtype=torch.float)
        x0, y0 = (cw - tw) // 2, (ch - fh) // 2
        canvas[:, y0:y0 + fh, x0:x0 + tw] = x
        h = int(np.clip(int(np.random.uniform(*self.scale) * fh), 1, ch))
        w = int(np.clip(int(np.random.uniform(*self.scale) * tw), 1, cw))
        i = random.randint(0, ch - h) if ch > h else 0
        j = random.randint(0, cw - w) if cw > w else 0
        crop = canvas[:, i:i + h, j:j + w]
        return F.interpolate(crop.unsqueeze(0), size=self.out_size, mode="bicubic", align_corners=True).squeeze(0)

    @staticmethod
    def _fade(x):
        head, tail = 2.0 * np.random.rand(2) - 1.0
        T = x.shape[2]
        return x + torch.linspace(head, tail, T, dtype=x.dtype).reshape(1, 1, T)

    def __call__(self, x: torch.Tensor) -> List[torch.Tensor]:
        return [self._fade(self._rrc(self._mix(x))) for _ in range(2)]


def clip_lms_path(lms_full: torch.Tensor, norm_stats: Sequence[float], tfm: PortPairTransform, crop_frames: int = 96):
    """AudioSet.__getitem__ arithmetic (datasets.py:336-357) on a precomputed (64, T_full) log-mel."""
    lms = lms_full.unsqueeze(0)
    l = lms.shape[-1]
    if l > crop_frames:
        start = np.random.randint(l - crop_frames)
        lms = lms[..., start:start + crop_frames]
    elif l < crop_frames:
        lms = F.pad(lms, (0, crop_frames - l), mode="constant", value=0)
    lms = (lms.to(torch.float) - norm_stats[0]) / norm_stats[1]
    return tfm(lms)


class PortBarlowTwinsLoss(nn.Module):
    """utils/loss.py:8-30 restated (single process)."""

    def __init__(self, dim: int, alpha=1.0, lmbda=0.005, hsic=False):
        super().__init__()
        self.bn = nn.BatchNorm1d(dim, affine=False)
        self.alpha, self.lmbda, self.hsic = alpha, lmbda, hsic

    def forward(self, z1, z2):
        c = self.bn(z1).T @ self.bn(z2)
        c.div_(z1.shape[0])
        on = torch.diagonal(c).add_(-1).pow_(2).sum()
        n = c.shape[0]
        off_el = c.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()
        off = off_el.add_(1).pow_(2).sum() if self.hsic else off_el.pow_(2).sum()
        return self.alpha * on + self.lmbda * off


# ------------------------------------------------------------------------------------------------
# timing helpers used by bench.py
# ------------------------------------------------------------------------------------------------
def _frontend_worker(args):
    """One DataLoader-worker-like process: full 10 s log-mel per clip + crop + normalise + two views."""
    seed, n_clips, n_samples, norm_stats = args
    import time
    torch.set_num_threads(1)
    np.random.seed(seed)
    random.seed(seed)
    g = torch.Generator().manual_seed(seed)
    wav = torch.clamp(0.1 * torch.randn(n_clips, n_samples, generator=g), -1, 1)
    mel = make_melspec()
    tfm = PortPairTransform()
    t0 = time.perf_counter()
    for b in range(n_clips):
        lms = log_mel(mel, wav[b])
        clip_lms_path(lms, norm_stats, tfm)
    return time.perf_counter() - t0


def time_frontend(n_clips_total: int, n_samples: int, workers: int, norm_stats=(-0.8294, 4.6230)) -> float:
    """Clips/s of the reference's execution model: `workers` single-threaded processes, each replaying
    Dataset.__getitem__ (wav -> log-mel -> crop -> z-score -> two views) on its share of the clips."""
    import multiprocessing as mp
    per = max(1, n_clips_total // workers)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        times = pool.map(_frontend_worker, [(1000 + w, per, n_samples, norm_stats) for w in range(workers)])
    return per * workers / max(times)


def time_loss(n: int, d: int, repeats: int = 1) -> float:
    """Seconds per fwd+bwd of the reference loss in fp32 on the CPU with all intra-op threads."""
    import time
    g = torch.Generator().manual_seed(1)
    z1 = torch.randn(n, d, generator=g).requires_grad_(True)
    z2 = (0.6 * z1.detach() + 0.8 * torch.randn(n, d, generator=g)).requires_grad_(True)
    crit = PortBarlowTwinsLoss(d)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        loss = crit(z1, z2)
        loss.backward()
        best = min(best, time.perf_counter() - t0)
        z1.grad = None
        z2.grad = None
    return best


def time_frontend_batched(n_clips: int, n_samples: int, norm_stats=(-0.8294, 4.6230)) -> float:
    """Clips/s of the single-process variant: ONE batched torchaudio MelSpectrogram call over
    the whole (B, L) batch with all intra-op threads, then the reference's per-sample crop / z-score / two-view loop."""
    import time
    np.random.seed(2000)
    random.seed(2000)
    g = torch.Generator().manual_seed(2000)
    wav = torch.clamp(0.1 * torch.randn(n_clips, n_samples, generator=g), -1, 1)
    mel = make_melspec()
    tfm = PortPairTransform()
    t0 = time.perf_counter()
    lms = log_mel(mel, wav)
    for b in range(n_clips):
        clip_lms_path(lms[b], norm_stats, tfm)
    return n_clips / (time.perf_counter() - t0)


def host_description() -> dict:
    """CPU model string, core / thread counts and library versions (printed beside every CPU number)."""
    import os
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    try:
        import torchaudio
        ta = torchaudio.__version__
    except Exception:
        ta = "unavailable"
    return {"cpu_model": model, "os_cpu_count": os.cpu_count(), "torch_num_threads": torch.get_num_threads(),
            "torch": torch.__version__, "torchaudio": ta}
