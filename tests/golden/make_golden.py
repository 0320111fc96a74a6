"""Generate the golden fixtures in this directory FROM THE LIVE REFERENCE.

Runs only in the build container (needs /root/reference and torchaudio); the fixtures it
writes (`*.npz`) are committed so that the tests and the GPU box never need the reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything here calls the reference's own modules (utils.loss.BarlowTwinsLoss,
augmentations.*, utils.transforms.AudioPairTransform) and the torchaudio transform the
reference constructs at datasets.py:39-48.  Draws are recorded by wrapping the RNG entry
points the reference calls, so the fixtures pin the *indices* (bank index, crop box) as well
as the floating-point outputs.
"""
import os
import random
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import torchaudio.transforms as AT  # noqa: E402
import augmentations  # noqa: E402  (reference)
import utils.loss as ref_loss  # noqa: E402  (reference)
import utils.transforms as ref_transforms  # noqa: E402  (reference)
from oracle.abt_oracle import synth_wave, synth_embeddings  # noqa: E402  (inputs only)

torch.set_num_threads(1)


def args_ns(**kw):
    base = dict(mixup=True, Gnoise=False, RRC=True, RLF=True, n_mels=64, crop_frames=96,
                virtual_crop_scale=[1.0, 1.5], local_crops_number=0, local_crops_size=[16, 16],
                projector_out_dim=64, HSIC=False, alpha=1.0, lmbda=0.005)
    base.update(kw)
    return types.SimpleNamespace(**base)


def gen_logmel():
    out = {}
    mel = AT.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160,
                            n_mels=64, f_min=60, f_max=7800, power=2)
    wav = synth_wave(3, 15200, seed=0)
    wav[2] *= 1e-3                       # quiet clip: exercises the +eps floor region
    wav[2, 5000:] = 0.0                  # exact digital silence -> log(eps)
    lms = (mel(torch.from_numpy(wav)) + torch.finfo().eps).log()
    out["wav"] = wav
    out["lms"] = lms.numpy()
    out["fb"] = mel.mel_scale.fb.numpy()
    # long clip, only a few frames kept (frame indices spread over the clip incl. both edges)
    wav10 = synth_wave(1, 160000, seed=5)
    lms10 = (mel(torch.from_numpy(wav10)) + torch.finfo().eps).log().numpy()
    keep = np.array([0, 1, 2, 3, 4, 500, 501, 996, 997, 998, 999, 1000])
    out["wav10_seed"] = np.array(5)
    out["lms10_frames"] = keep
    out["lms10_cols"] = lms10[0][:, keep]
    out["lms10_shape"] = np.array(lms10.shape)
    # HEAR-style short window (hear/config.yaml: win_length 400)
    mel400 = AT.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=400, hop_length=160,
                               n_mels=64, f_min=60, f_max=7800, power=2)
    out["lms_win400"] = (mel400(torch.from_numpy(wav[:1])) + torch.finfo().eps).log().numpy()
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **out)


class _RecordingRandom:
    """Stands in for `np.random` inside the reference's augmentations module."""

    def __init__(self, log):
        self._log = log

    def random(self, *a):
        v = np.random.random(*a)
        self._log.append(("np.random", float(v)))
        return v

    def randint(self, *a, **k):
        v = np.random.randint(*a, **k)
        self._log.append(("np.randint", int(v)))
        return v

    def uniform(self, *a):
        v = np.random.uniform(*a)
        self._log.append(("np.uniform", float(v)))
        return v

    def rand(self, *a):
        v = np.random.rand(*a)
        self._log.append(("np.rand", [float(x) for x in np.atleast_1d(v)]))
        return v


def gen_views():
    out = {}
    log = []
    np_proxy = types.SimpleNamespace(**{k: getattr(np, k) for k in ("clip",)})
    np_proxy.random = _RecordingRandom(log)
    orig_np = augmentations.np
    orig_get = augmentations.RandomResizeCrop.get_params
    boxes = []

    def rec_get(*a):
        r = orig_get(*a)
        boxes.append([int(x) for x in r])
        return r

    augmentations.np = np_proxy
    augmentations.RandomResizeCrop.get_params = staticmethod(rec_get)
    try:
        seed = 1234
        np.random.seed(seed)
        random.seed(seed)
        B = 6
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, 1, 64, 96, generator=g) * 1.1 + 0.1
        tfm = ref_transforms.AudioPairTransform(args_ns())
        views = []
        for b in range(B):
            crops = tfm(x[b])
            views.append(torch.stack(crops).numpy())
        out["seed"] = np.array(seed)
        out["x"] = x.numpy()
        out["views"] = np.stack(views)                      # (B, 2, 1, 64, 96)
        out["boxes"] = np.array(boxes, dtype=np.int64)      # (B*2, 4) i,j,h,w
        alphas, bank_idx, fades = [], [], []
        for kind, v in log:
            if kind == "np.random":
                alphas.append(v)
            elif kind == "np.randint":
                bank_idx.append(v)
            elif kind == "np.rand":
                fades.append(v)
        out["mix_u"] = np.array(alphas)                      # raw U(0,1) draws, alpha = 0.2*u
        out["bank_idx"] = np.array(bank_idx, dtype=np.int64)  # one per view except the very first
        out["fade_u"] = np.array(fades)                      # (B*2, 2) raw U(0,1) draws

        # lms path of AudioSet.__getitem__ (datasets.py:336-357) on a long log-mel with the time crop
        log.clear(); boxes.clear()
        np.random.seed(77); random.seed(77)
        lms_full = (torch.randn(3, 64, 301, generator=g) * 4.0 - 1.0)
        tfm2 = ref_transforms.AudioPairTransform(args_ns())
        starts, v2, xs = [], [], []
        for b in range(3):
            lms = lms_full[b].unsqueeze(0)
            l = lms.shape[-1]
            start = np.random.randint(l - 96)
            starts.append(int(start))
            lms = lms[..., start:start + 96].to(torch.float)
            lms = (lms - (-0.8294)) / 4.6230
            xs.append(lms.numpy())
            v2.append(torch.stack(tfm2(lms)).numpy())
        out["lms_full"] = lms_full.numpy()
        out["lms_starts"] = np.array(starts, dtype=np.int64)
        out["lms_x"] = np.stack(xs)
        out["lms_views"] = np.stack(v2)
        out["lms_boxes"] = np.array(boxes, dtype=np.int64)

        # multi-crop with 2 local crops, no mixup
        log.clear(); boxes.clear()
        np.random.seed(5); random.seed(5)
        tfm3 = ref_transforms.AudioPairTransform(args_ns(mixup=False, local_crops_number=2))
        crops = tfm3(x[0])
        out["mc_global"] = torch.stack(crops[:2]).numpy()
        out["mc_local"] = torch.stack(crops[2:]).numpy()
        out["mc_boxes"] = np.array(boxes, dtype=np.int64)
    finally:
        augmentations.np = orig_np
        augmentations.RandomResizeCrop.get_params = orig_get

    # plain bicubic known answers (F.interpolate as called at augmentations.py:53-54)
    g = torch.Generator().manual_seed(3)
    cases = [(64, 96), (38, 57), (64, 143), (45, 120), (1, 1), (3, 2)]
    for k, (h, w) in enumerate(cases):
        src = torch.randn(1, 1, h, w, generator=g)
        dst = torch.nn.functional.interpolate(src, size=(64, 96), mode="bicubic", align_corners=True)
        out[f"bic_src{k}"] = src[0, 0].numpy()
        out[f"bic_dst{k}"] = dst[0, 0].numpy()
    out["fader_lin"] = torch.linspace(-0.37, 0.81, 96, dtype=torch.float32).numpy()
    np.savez_compressed(os.path.join(HERE, "views.npz"), **out)


def gen_loss():
    out = {}
    for tag, (n, d, hsic) in {"a": (32, 64, False), "b": (16, 128, True), "c": (48, 256, False)}.items():
        z1, z2 = synth_embeddings(n, d, seed=11 + n)
        cfg = args_ns(projector_out_dim=d, HSIC=hsic, alpha=1.0, lmbda=0.005)
        mod = ref_loss.BarlowTwinsLoss(cfg, ncrops=2)
        t1 = torch.from_numpy(z1).requires_grad_(True)
        t2 = torch.from_numpy(z2).requires_grad_(True)
        loss = mod(t2, t1, ngcrops_each=1)      # forward(student, teacher): forward_loss(teacher, student)
        loss.backward()
        out[f"{tag}_z1"] = z1
        out[f"{tag}_z2"] = z2
        out[f"{tag}_hsic"] = np.array(hsic)
        out[f"{tag}_loss"] = loss.detach().numpy()
        out[f"{tag}_dz1"] = t1.grad.numpy()
        out[f"{tag}_dz2"] = t2.grad.numpy()
        out[f"{tag}_running_mean"] = mod.bn.running_mean.numpy()
        out[f"{tag}_running_var"] = mod.bn.running_var.numpy()
        out[f"{tag}_num_batches"] = mod.bn.num_batches_tracked.numpy()
    # main_bt_byol.py pairing: ngcrops_each=2, ncrops=2 -> two terms
    n, d = 24, 64
    za, zb = synth_embeddings(2 * n, d, seed=99)
    cfg = args_ns(projector_out_dim=d)
    mod = ref_loss.BarlowTwinsLoss(cfg, ncrops=2)
    ts = torch.from_numpy(za).requires_grad_(True)
    tt = torch.from_numpy(zb).requires_grad_(True)
    loss = mod(ts, tt, ngcrops_each=2)
    loss.backward()
    out["byol_student"] = za
    out["byol_teacher"] = zb
    out["byol_loss"] = loss.detach().numpy()
    out["byol_dstudent"] = ts.grad.numpy()
    out["byol_dteacher"] = tt.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


if __name__ == "__main__":
    gen_logmel()
    gen_views()
    gen_loss()
    for f in ("logmel.npz", "views.npz", "loss.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
