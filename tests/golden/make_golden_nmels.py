"""Golden log-mels for band counts other than the reference default (hyperparameters.py exposes --n_mels): torchaudio's own
MelSpectrogram, as datasets.py:39-48,115 builds it.  Runs only in the build container (needs torchaudio); the fixture is committed.
    python tests/golden/make_golden_nmels.py"""
import os
import sys

import numpy as np
import torch
import torchaudio.transforms as AT

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.abt_oracle import synth_wave  # noqa: E402

out = {}
wav = synth_wave(2, 12000, seed=11)
out["wav"] = wav
for n_mels, f_min, f_max in ((80, 60, 7800), (128, 0, 8000), (40, 60, 7800)):
    mel = AT.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=160, n_mels=n_mels, f_min=f_min, f_max=f_max, power=2)
    out[f"lms_{n_mels}"] = (mel(torch.from_numpy(wav)) + torch.finfo().eps).log().numpy()
    out[f"cfg_{n_mels}"] = np.array([n_mels, f_min, f_max])
np.savez_compressed(os.path.join(HERE, "logmel_nmels.npz"), **out)
print("wrote logmel_nmels.npz", {k: v.shape for k, v in out.items()})
