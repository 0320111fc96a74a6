"""CPU oracle for the Audio Barlow Twins hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a numpy restatement of the arithmetic the reference (jonahanton/SSL_audio,
mounted at /root/reference when it exists) performs on the path named by
BASELINE.json:north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``ssl_audio_b200`` never does and fails loudly without its CUDA library.

Parity pin: the reference ships no golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference and torchaudio) and committed
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here against
those fixtures.

Each function cites the reference lines it restates (paths relative to /root/reference;
``torchaudio/...`` = the installed torchaudio 2.11 wheel the reference calls into).
"""
from __future__ import annotations

import math
import random as _pyrandom
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32_EPS = float(np.finfo(np.float32).eps)  # torch.finfo().eps, datasets.py:115


# --------------------------------------------------------------------------------------
# Frontend: wav -> power STFT -> mel -> log -> normalise
# --------------------------------------------------------------------------------------

def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace(dtype=float32) semantics: step = (end-start)/(steps-1) in fp32; the first
    half is evaluated from `start`, the second half from `end`, each with ONE rounding (ATen's
    vectorised kernel fuses the multiply-add; emulated here in float64, where the fp32*int
    product is exact).  Bit-exact against torch 2.11 (tests/golden/views.npz:fader_lin).  Used by
    the reference at augmentations.py:72 and inside torchaudio/functional/functional.py:559,565."""
    start32, end32 = np.float32(start), np.float32(end)
    if steps == 1:
        return np.array([start32], dtype=np.float32)
    step = np.float64(np.float32((end32 - start32) / np.float32(steps - 1)))
    idx = np.arange(steps)
    half = steps // 2
    lo = (np.float64(start32) + step * idx).astype(np.float32)
    hi = (np.float64(end32) - step * (steps - 1 - idx)).astype(np.float32)
    return np.where(idx < half, lo, hi).astype(np.float32)


def hann_periodic(win_length: int) -> np.ndarray:
    """torch.hann_window(win_length) (periodic) as used by torchaudio Spectrogram
    (torchaudio/transforms/_transforms.py Spectrogram.__init__, called from datasets.py:39)."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)).astype(np.float32)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """Triangular HTK mel filterbank, norm=None: torchaudio/functional/functional.py:518-587
    (+ _hz_to_mel :425-445, _mel_to_hz :459-474, _create_triangular_filterbank :492-515).
    Returns (n_freqs, n_mels) float32."""
    all_freqs = linspace_f32(0.0, float(sample_rate // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = linspace_f32(m_min, m_max, n_mels + 2)
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0)) - np.float32(1.0))).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


@dataclass
class MelConfig:
    """Fields of `args` read at datasets.py:39-48 (defaults utils/hyperparameters.py:42-60)."""
    sample_rate: int = 16000
    n_fft: int = 1024
    win_length: int = 1024
    hop_length: int = 160
    n_mels: int = 64
    f_min: float = 60.0
    f_max: float = 7800.0


def power_stft(wav: np.ndarray, cfg: MelConfig) -> np.ndarray:
    """|STFT|^2 as torchaudio Spectrogram(power=2): torchaudio/functional/functional.py:123-144
    -> torch.stft(center=True, pad_mode='reflect', onesided=True, window=hann periodic).
    wav (..., L) -> (..., n_fft//2+1, 1 + L//hop).  The window (win_length <= n_fft) is
    centre-padded to n_fft as torch.stft does."""
    wav = np.asarray(wav, dtype=np.float32)
    lead = wav.shape[:-1]
    x = wav.reshape(-1, wav.shape[-1])
    pad = cfg.n_fft // 2
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = 1 + x.shape[-1] // cfg.hop_length
    win = hann_periodic(cfg.win_length)
    if cfg.win_length < cfg.n_fft:
        left = (cfg.n_fft - cfg.win_length) // 2
        win = np.pad(win, (left, cfg.n_fft - cfg.win_length - left))
    idx = np.arange(n_frames)[:, None] * cfg.hop_length + np.arange(cfg.n_fft)[None, :]
    frames = xp[:, idx] * win[None, None, :]                       # (B, T, n_fft)
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
    power = (spec.real ** 2 + spec.imag ** 2)                        # float64
    return np.swapaxes(power, -1, -2).reshape(*lead, cfg.n_fft // 2 + 1, n_frames)


def log_mel(wav: np.ndarray, cfg: MelConfig = MelConfig()) -> np.ndarray:
    """(MelSpectrogram(wav) + eps).log(): datasets.py:115, old/data_manager/wav_to_lms.py:58-61;
    MelScale = matmul(spec^T, fb)^T at torchaudio/transforms/_transforms.py:407-419.
    wav (..., L) -> (..., n_mels, 1 + L//hop) float32."""
    power = power_stft(wav, cfg)
    fb = melscale_fbanks_htk(cfg.n_fft // 2 + 1, cfg.f_min, cfg.f_max, cfg.n_mels, cfg.sample_rate)
    mel = np.einsum("...kt,km->...mt", power, fb.astype(np.float64))
    return np.log(mel.astype(np.float32) + np.float32(F32_EPS)).astype(np.float32)


def wav_unit_pad(wav: np.ndarray, unit_length: int) -> np.ndarray:
    """Centre zero-pad a short waveform to unit_length: datasets.py:103-108."""
    adj = unit_length - len(wav)
    if adj > 0:
        half = adj // 2
        wav = np.pad(wav, (half, adj - half))
    return wav


def wav_unit_crop(wav: np.ndarray, unit_length: int) -> Tuple[np.ndarray, int]:
    """Random unit-length crop with CPython `random.randint`: datasets.py:110-113."""
    adj = len(wav) - unit_length
    start = _pyrandom.randint(0, adj) if adj > 0 else 0
    return wav[start:start + unit_length], start


def lms_trim_pad(lms: np.ndarray, crop_frames: int) -> Tuple[np.ndarray, int]:
    """Random time crop (np.random.randint, exclusive high) or right zero-pad:
    datasets.py:87-96 and :342-351.  lms (..., F, T).  Returns (lms, start) with start=-1
    when no crop was drawn."""
    l = lms.shape[-1]
    start = -1
    if l > crop_frames:
        start = int(np.random.randint(l - crop_frames))
        lms = lms[..., start:start + crop_frames]
    elif l < crop_frames:
        padw = [(0, 0)] * (lms.ndim - 1) + [(0, crop_frames - l)]
        lms = np.pad(lms, padw, mode="constant", constant_values=0.0)
    return lms.astype(np.float32), start


def normalise(lms: np.ndarray, norm_stats: Sequence[float]) -> np.ndarray:
    """(lms - mean) / std with python-float stats against an fp32 tensor: datasets.py:118-119."""
    return ((lms.astype(np.float32) - np.float32(norm_stats[0])) / np.float32(norm_stats[1])).astype(np.float32)


# --------------------------------------------------------------------------------------
# Augmentations
# --------------------------------------------------------------------------------------

def log_mixup_exp(xa: np.ndarray, xb: np.ndarray, alpha: float) -> np.ndarray:
    """augmentations.py:81-85.  `alpha` is a python float; torch multiplies it into fp32
    tensors, i.e. the weights are float32(alpha) and float32(1. - alpha)."""
    wa = np.float32(alpha)
    wb = np.float32(1.0 - alpha)
    x = wa * np.exp(xa.astype(np.float32)) + wb * np.exp(xb.astype(np.float32))
    return np.log(x + np.float32(F32_EPS)).astype(np.float32)


@dataclass
class MixupState:
    """MixupBYOLA state: augmentations.py:96-101.  `bank` holds the un-mixed inputs."""
    ratio: float = 0.2
    n_memory: int = 2048
    bank: List[np.ndarray] = field(default_factory=list)


def mixup_byola(x: np.ndarray, st: MixupState) -> Tuple[np.ndarray, dict]:
    """MixupBYOLA.forward: augmentations.py:103-117.  Returns (mixed, params) where params
    records alpha and the bank index drawn (-1 when the bank was empty)."""
    alpha = st.ratio * np.random.random()
    idx = -1
    if st.bank:
        idx = int(np.random.randint(len(st.bank)))
        mixed = log_mixup_exp(x, st.bank[idx], 1.0 - alpha)
    else:
        mixed = x
    st.bank = (st.bank + [x])[-st.n_memory:]
    return mixed.astype(np.float32), {"alpha": alpha, "bank_index": idx, "bank_len_before": len(st.bank) - 1 if idx < 0 else None}


def mix_gaussian_noise(lms: np.ndarray, ratio: float, n: np.ndarray) -> Tuple[np.ndarray, float]:
    """MixGaussianNoise.forward (augmentations.py:133-140): lambd = ratio * np.random.rand(); z = exp(normal(0, lambd));
    log((1 - lambd) * exp(lms) + z + eps).  `n` are the STANDARD normal draws behind torch.normal(0, lambd, shape) -- ATen fills
    N(0, 1) and then scales by float32(lambd) -- so passing the reference's own draws reproduces it; fp32 arithmetic throughout."""
    lambd = ratio * np.random.rand()
    x = np.exp(np.asarray(lms, dtype=np.float32))
    z = np.exp((np.asarray(n, dtype=np.float32) * np.float32(lambd)).astype(np.float32))
    mixed = np.float32(1 - lambd) * x + z + F32_EPS
    return np.log(mixed).astype(np.float32), lambd


def rrc_get_params(canvas_hw, in_hw, time_scale, freq_scale) -> Tuple[int, int, int, int]:
    """RandomResizeCrop.get_params: augmentations.py:30-38 (numpy uniform x2, then CPython
    random.randint only when the canvas is larger than the crop)."""
    canvas_h, canvas_w = canvas_hw
    src_h, src_w = in_hw
    h = int(np.clip(int(np.random.uniform(*freq_scale) * src_h), 1, canvas_h))
    w = int(np.clip(int(np.random.uniform(*time_scale) * src_w), 1, canvas_w))
    i = _pyrandom.randint(0, canvas_h - h) if canvas_h > h else 0
    j = _pyrandom.randint(0, canvas_w - w) if canvas_w > w else 0
    return i, j, h, w


def _cubic_coeffs(t: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """ATen get_cubic_upsample_coefficients (A = -0.75), fp32, as used by
    F.interpolate(mode='bicubic') at augmentations.py:53-54."""
    A = np.float32(-0.75)
    one = np.float32(1.0)

    def cc1(x):  # |x| <= 1
        return ((A + np.float32(2.0)) * x - (A + np.float32(3.0))) * x * x + one

    def cc2(x):  # 1 < |x| < 2
        return ((A * x - np.float32(5.0) * A) * x + np.float32(8.0) * A) * x - np.float32(4.0) * A

    return cc2(t + one), cc1(t), cc1(one - t), cc2(np.float32(2.0) - t)


def bicubic_resize_align_corners(src: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """F.interpolate(src[None,None], size=out_hw, mode='bicubic', align_corners=True):
    ATen upsample_bicubic2d; source index = dst * (in-1)/(out-1), 4x4 taps clamped to the
    border.  src (h, w) float32 -> (out_h, out_w) float32."""
    src = src.astype(np.float32)
    h, w = src.shape
    oh, ow = out_hw

    def axis(n_in, n_out):
        scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0.0)
        pos = scale * np.arange(n_out, dtype=np.float32)
        base = np.floor(pos)
        t = (pos - base).astype(np.float32)
        base = base.astype(np.int64)
        taps = np.stack([np.clip(base + k, 0, n_in - 1) for k in (-1, 0, 1, 2)], axis=0)  # (4, n_out)
        coef = np.stack(_cubic_coeffs(t), axis=0).astype(np.float32)                       # (4, n_out)
        return taps, coef

    ty, cy = axis(h, oh)
    tx, cx = axis(w, ow)
    # horizontal pass on each of the 4 tap rows, then vertical combination (ATen's order)
    out = np.zeros((oh, ow), dtype=np.float32)
    for a in range(4):
        rows = src[ty[a]]                                    # (oh, w)
        horiz = np.zeros((oh, ow), dtype=np.float32)
        for b in range(4):
            horiz += rows[:, tx[b]] * cx[b][None, :]
        out += horiz * cy[a][:, None]
    return out


def random_resize_crop(lms: np.ndarray, out_size=(64, 96), virtual_crop_scale=(1.0, 1.5),
                       freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5), params=None) -> Tuple[np.ndarray, Tuple[int, int, int, int]]:
    """RandomResizeCrop.forward: augmentations.py:40-55.  lms (1, F, T).  `params`=(i,j,h,w)
    replays recorded draws instead of consuming the RNG."""
    c, F, T = lms.shape
    ch, cw = int(F * virtual_crop_scale[0]), int(T * virtual_crop_scale[1])
    canvas = np.zeros((c, ch, cw), dtype=np.float32)
    x0, y0 = (cw - T) // 2, (ch - F) // 2
    canvas[:, y0:y0 + F, x0:x0 + T] = lms
    if params is None:
        # NB the reference passes (time_scale, freq_scale) positionally into a signature that
        # names them (time_scale, freq_scale): augmentations.py:31,50 -- no swap.
        params = rrc_get_params((ch, cw), (F, T), time_scale, freq_scale)
    i, j, h, w = params
    crop = canvas[:, i:i + h, j:j + w]
    out = np.stack([bicubic_resize_align_corners(crop[k], out_size) for k in range(c)], axis=0)
    return out.astype(np.float32), (i, j, h, w)


def linear_fader(lms: np.ndarray, gain: float = 1.0, params=None) -> Tuple[np.ndarray, Tuple[float, float]]:
    """RandomLinearFader.forward: augmentations.py:69-74."""
    if params is None:
        head, tail = gain * ((2.0 * np.random.rand(2)) - 1.0)
    else:
        head, tail = params
    T = lms.shape[2]
    slope = linspace_f32(float(head), float(tail), T).reshape(1, 1, T)
    return (lms.astype(np.float32) + slope).astype(np.float32), (float(head), float(tail))


@dataclass
class PairTransformConfig:
    """Fields of `args` read by AudioPairTransform: utils/transforms.py:14-45."""
    mixup: bool = True
    Gnoise: bool = False
    RRC: bool = True
    RLF: bool = True
    n_mels: int = 64
    crop_frames: int = 96
    virtual_crop_scale: Tuple[float, float] = (1.0, 1.5)
    local_crops_number: int = 0
    local_crops_size: Tuple[int, int] = (16, 16)
    mixup_ratio: float = 0.2
    gauss_noise_ratio: float = 0.2
    global_crop_scale: Tuple[float, float] = (0.6, 1.5)
    local_crop_scale: Tuple[float, float] = (0.05, 0.6)


def global_view(x: np.ndarray, cfg: PairTransformConfig, st: MixupState, noise: Optional[np.ndarray] = None) -> Tuple[np.ndarray, dict]:
    """One pass of AudioPairTransform.global_transform: utils/transforms.py:16-34 (Mixup -> Gnoise -> RRC -> RLF).
    `noise`: the N(0, 1) draws of MixGaussianNoise for this view (required with cfg.Gnoise)."""
    rec = {}
    y = x
    if cfg.mixup:
        y, p = mixup_byola(y, st)
        rec.update(alpha=p["alpha"], bank_index=p["bank_index"])
    if cfg.Gnoise:
        if noise is None:
            raise ValueError("cfg.Gnoise needs the standard-normal draws of torch.normal (augmentations.py:137)")
        y, lambd = mix_gaussian_noise(y, cfg.gauss_noise_ratio, noise)
        rec.update(lambd=lambd)
    if cfg.RRC:
        y, (i, j, h, w) = random_resize_crop(y, (cfg.n_mels, cfg.crop_frames), tuple(cfg.virtual_crop_scale),
                                             cfg.global_crop_scale, cfg.global_crop_scale)
        rec.update(i=i, j=j, h=h, w=w)
    if cfg.RLF:
        y, (head, tail) = linear_fader(y)
        rec.update(head=head, tail=tail)
    return y, rec


def local_view(x: np.ndarray, cfg: PairTransformConfig) -> Tuple[np.ndarray, dict]:
    """AudioPairTransform.local_transform: utils/transforms.py:37-47 (RRC to local_crops_size,
    virtual_crop_scale (1,1), scale local_crop_scale)."""
    y, (i, j, h, w) = random_resize_crop(x, tuple(cfg.local_crops_size), (1, 1), cfg.local_crop_scale, cfg.local_crop_scale)
    return y, dict(i=i, j=j, h=h, w=w)


def audio_pair_transform(x: np.ndarray, cfg: PairTransformConfig, st: MixupState,
                         noise: Optional[np.ndarray] = None) -> Tuple[List[np.ndarray], List[dict]]:
    """AudioPairTransform.forward (multi_transform=True): utils/transforms.py:49-58.  noise: (2, 1, F, T) with cfg.Gnoise."""
    outs, recs = [], []
    for k in range(2):
        y, r = global_view(x, cfg, st, None if noise is None else noise[k])
        outs.append(y)
        recs.append(r)
    for _ in range(cfg.local_crops_number):
        y, r = local_view(x, cfg)
        outs.append(y)
        recs.append(r)
    return outs, recs


def frontend_clip_lms_path(lms_full: np.ndarray, norm_stats, cfg: PairTransformConfig, st: MixupState):
    """AudioSet.__getitem__ arithmetic on a precomputed log-mel (64, T_full): datasets.py:336-357."""
    lms = lms_full[None].astype(np.float32)
    lms, start = lms_trim_pad(lms, cfg.crop_frames)
    if norm_stats is not None:
        lms = normalise(lms, norm_stats)
    views, recs = audio_pair_transform(lms, cfg, st)
    return views, dict(start=start, views=recs), lms


def frontend_clip_wav_path(wav: np.ndarray, mel_cfg: MelConfig, unit_sec: float, norm_stats,
                           cfg: PairTransformConfig, st: MixupState):
    """FSD50K.__getitem__ arithmetic on a raw waveform: datasets.py:98-122."""
    unit_length = int(unit_sec * mel_cfg.sample_rate)
    wav = wav_unit_pad(np.asarray(wav, dtype=np.float32), unit_length)
    wav, start = wav_unit_crop(wav, unit_length)
    lms = log_mel(wav, mel_cfg)[None]
    if norm_stats is not None:
        lms = normalise(lms, norm_stats)
    views, recs = audio_pair_transform(lms, cfg, st)
    return views, dict(start=start, views=recs), lms


def normalize_batch(x: np.ndarray) -> np.ndarray:
    """NormalizeBatch.forward (augmentations.py:228-231) with its default axis [0, 2, 3]: per-channel mean and UNBIASED std
    (torch.std default) over batch, frequency and time, std clamped to [finfo.eps, finfo.max].  x: (B, C, F, T) float32."""
    x = np.asarray(x, dtype=np.float32)
    mean = x.astype(np.float64).mean(axis=(0, 2, 3), keepdims=True)
    std = x.astype(np.float64).std(axis=(0, 2, 3), ddof=1, keepdims=True)
    std = np.clip(std, F32_EPS, np.finfo(np.float32).max)
    return ((x - mean) / std).astype(np.float32)


class RunningNormState:
    """RunningNorm (augmentations.py:187-210) with its default axis [1, 2] on (1, F, T) samples: scalar running mean of the sample
    means and of mean((x - mu)^2), both with the reference's RunningMean update `mu += (m - mu) / n` where n is the number of samples
    seen BEFORE this one (augmentations.py:150-156) -- not the textbook 1/(n+1)."""

    def __init__(self, epoch_samples: int, max_update_epochs: int = 10):
        self.max_update = epoch_samples * max_update_epochs
        self.n = 0
        self.mu = 0.0
        self.s2 = 0.0

    def put(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float32)
        if self.n < self.max_update:
            m = float(x.astype(np.float64).mean())
            self.mu = m if self.n == 0 else self.mu + (m - self.mu) / self.n
            v = float(((x.astype(np.float64) - self.mu) ** 2).mean())
            self.s2 = v if self.n == 0 else self.s2 + (v - self.s2) / self.n
            self.n += 1
        std = min(max(np.float32(np.sqrt(self.s2)), F32_EPS), np.finfo(np.float32).max)
        return ((x - np.float32(self.mu)) / np.float32(std)).astype(np.float32)


def running_norm(samples: np.ndarray, epoch_samples: int, max_update_epochs: int = 10) -> np.ndarray:
    """A (B, 1, F, T) batch pushed through RunningNormState in sample order."""
    st = RunningNormState(epoch_samples, max_update_epochs)
    return np.stack([st.put(x) for x in samples])


# --------------------------------------------------------------------------------------
# Barlow Twins objective
# --------------------------------------------------------------------------------------

def off_diagonal(x: np.ndarray) -> np.ndarray:
    """utils/utils.py:23-27."""
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].reshape(n - 1, n + 1)[:, 1:].flatten()


def batchnorm_train(z: np.ndarray, eps: float = 1e-5):
    """nn.BatchNorm1d(D, affine=False) in training mode (utils/loss.py:13,17): biased variance."""
    mu = z.mean(axis=0)
    var = z.var(axis=0)
    rstd = 1.0 / np.sqrt(var + eps)
    return (z - mu) * rstd, mu, var, rstd


def bt_loss_forward(z1: np.ndarray, z2: np.ndarray, alpha=1.0, lmbda=0.005, hsic=False, eps=1e-5,
                    dtype=np.float64):
    """BarlowTwinsLoss.forward_loss (single process): utils/loss.py:15-30."""
    z1 = z1.astype(dtype)
    z2 = z2.astype(dtype)
    n = z1.shape[0]
    h1, _, _, _ = batchnorm_train(z1, eps)
    h2, _, _, _ = batchnorm_train(z2, eps)
    c = h1.T @ h2 / n
    on = ((np.diagonal(c) - 1.0) ** 2).sum()
    off_el = off_diagonal(c)
    off = ((off_el + 1.0) ** 2).sum() if hsic else (off_el ** 2).sum()
    return alpha * on + lmbda * off, c


def bt_loss_forward_backward(z1: np.ndarray, z2: np.ndarray, alpha=1.0, lmbda=0.005, hsic=False, eps=1e-5,
                             dtype=np.float64):
    """Loss and d loss / d z1, d z2 in closed form: the autograd graph of utils/loss.py:15-30
    (matmul backward + batch_norm backward).  Returns (loss, dz1, dz2, c)."""
    z1 = z1.astype(dtype)
    z2 = z2.astype(dtype)
    n = z1.shape[0]
    h1, _, _, r1 = batchnorm_train(z1, eps)
    h2, _, _, r2 = batchnorm_train(z2, eps)
    c = h1.T @ h2 / n
    d = c.shape[0]
    on = ((np.diagonal(c) - 1.0) ** 2).sum()
    shift = 1.0 if hsic else 0.0
    g = 2.0 * lmbda * (c + shift)
    g[np.arange(d), np.arange(d)] = 2.0 * alpha * (np.diagonal(c) - 1.0)
    off_el = off_diagonal(c) + shift
    loss = alpha * on + lmbda * (off_el ** 2).sum()
    gh1 = h2 @ g.T / n
    gh2 = h1 @ g / n

    def bn_bwd(gh, h, r):
        return (gh - gh.mean(axis=0) - h * (gh * h).mean(axis=0)) * r

    return loss, bn_bwd(gh1, h1, r1), bn_bwd(gh2, h2, r2), c


def bt_loss_forward_backward_blocked(z1: np.ndarray, z2: np.ndarray, alpha=1.0, lmbda=0.005, hsic=False, eps=1e-5,
                                     block: int = 1024, dtype=np.float64):
    """Same closed form as bt_loss_forward_backward (utils/loss.py:15-30 + its autograd graph), evaluated one
    row block of C at a time so that D = 8192 never holds a D x D float64 matrix.  Returns (loss, dz1, dz2)."""
    z1 = z1.astype(dtype)
    z2 = z2.astype(dtype)
    n, d = z1.shape
    h1, _, _, r1 = batchnorm_train(z1, eps)
    h2, _, _, r2 = batchnorm_train(z2, eps)
    shift = 1.0 if hsic else 0.0
    on = 0.0
    off = 0.0
    gh1 = np.empty_like(h1)
    gh2 = np.zeros_like(h2)
    for s in range(0, d, block):
        e = min(s + block, d)
        c = h1[:, s:e].T @ h2 / n                      # rows s..e of C
        idx = np.arange(s, e)
        dg = c[idx - s, idx].copy()
        on += ((dg - 1.0) ** 2).sum()
        cs = c + shift
        off += (cs ** 2).sum() - ((dg + shift) ** 2).sum()
        g = 2.0 * lmbda * cs
        g[idx - s, idx] = 2.0 * alpha * (dg - 1.0)
        gh1[:, s:e] = h2 @ g.T / n
        gh2 += h1[:, s:e] @ g / n

    def bn_bwd(gh, h, r):
        return (gh - gh.mean(axis=0) - h * (gh * h).mean(axis=0)) * r

    return alpha * on + lmbda * off, bn_bwd(gh1, h1, r1), bn_bwd(gh2, h2, r2)


def bn_running_update(running_mean, running_var, z, momentum=0.1):
    """BatchNorm1d running-stat update (unbiased variance), applied z1 then z2: utils/loss.py:17."""
    n = z.shape[0]
    mu = z.astype(np.float64).mean(axis=0)
    var_unbiased = z.astype(np.float64).var(axis=0) * (n / max(n - 1, 1))
    rm = (1 - momentum) * running_mean + momentum * mu
    rv = (1 - momentum) * running_var + momentum * var_unbiased
    return rm, rv


def bt_loss_multicrop(student: np.ndarray, teacher: np.ndarray, ncrops: int, ngcrops_each: int = 1, **kw):
    """BarlowTwinsLoss.forward pairing loop: utils/loss.py:32-48.  Returns (loss, dstudent, dteacher)."""
    s_chunks = np.array_split(student, ncrops - (2 - ngcrops_each), axis=0)
    t_chunks = np.array_split(teacher, ngcrops_each, axis=0)
    total = 0.0
    ds = [np.zeros_like(c, dtype=np.float64) for c in s_chunks]
    dt = [np.zeros_like(c, dtype=np.float64) for c in t_chunks]
    terms = 0
    for q in range(len(t_chunks)):
        for v in range(len(s_chunks)):
            if len(t_chunks) > 1 and q == v:
                continue
            loss, dz1, dz2, _ = bt_loss_forward_backward(t_chunks[q], s_chunks[v], **kw)
            total += loss
            dt[q] += dz1
            ds[v] += dz2
            terms += 1
    return total / terms, np.concatenate(ds, 0) / terms, np.concatenate(dt, 0) / terms


# --------------------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------------------

def proj_tail_loss(h1: np.ndarray, h2: np.ndarray, w: np.ndarray, alpha=1.0, lmbda=0.005, hsic=False):
    """Last bias-free Linear of BarlowTwinsHead (model.py:22, 25-31) + forward_loss (utils/loss.py:15-30) with bf16 embeddings:
    z = round_bf16(h W^T) (float64 accumulation), then the objective on z.  Returns (loss, z1, z2, dz1, dz2, dh1, dh2, dW) with the
    Linear's backward in float64 (dh = dz W, dW = dz1^T h1 + dz2^T h2)."""
    h1 = h1.astype(np.float64); h2 = h2.astype(np.float64); w = w.astype(np.float64)
    z1 = round_bf16((h1 @ w.T).astype(np.float32))
    z2 = round_bf16((h2 @ w.T).astype(np.float32))
    loss, dz1, dz2, _ = bt_loss_forward_backward(z1, z2, alpha, lmbda, hsic)
    return loss, z1, z2, dz1, dz2, dz1 @ w, dz2 @ w, dz1.T @ h1 + dz2.T @ h2


def synth_wave(batch: int, length: int, seed: int = 0, sample_rate: int = 16000) -> np.ndarray:
    """0.1*white noise + 3 sinusoids (0.3/0.1/0.03, U(100,7000) Hz, random phase), clamped to
    [-1, 1], float32 (B, L)."""
    rng = np.random.default_rng(seed)
    t = np.arange(length, dtype=np.float64) / sample_rate
    out = 0.1 * rng.standard_normal((batch, length))
    for amp in (0.3, 0.1, 0.03):
        f = rng.uniform(100.0, 7000.0, size=(batch, 1))
        ph = rng.uniform(0.0, 2 * np.pi, size=(batch, 1))
        out += amp * np.sin(2 * np.pi * f * t[None, :] + ph)
    return np.clip(out, -1.0, 1.0).astype(np.float32)


def synth_embeddings(n: int, d: int, seed: int = 1):
    """z1 = randn, z2 = 0.6 z1 + 0.8 randn, both rounded to bf16-representable float32."""
    rng = np.random.default_rng(seed)
    z1 = rng.standard_normal((n, d)).astype(np.float32)
    z2 = (0.6 * z1 + 0.8 * rng.standard_normal((n, d))).astype(np.float32)
    return round_bf16(z1), round_bf16(z2)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)
