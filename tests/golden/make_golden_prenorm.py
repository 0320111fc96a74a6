"""Golden fixture for RunningNorm FROM THE LIVE REFERENCE (build container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_prenorm.py

Two sequences: "a" keeps updating over all 12 samples; "b" has epoch_samples=1, max_update_epochs=5, so the statistics freeze after
five samples and the remaining three are normalised with the frozen values.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import augmentations  # noqa: E402  (reference)

torch.set_num_threads(1)
g = torch.Generator().manual_seed(23)
out = {}
for tag, n, epoch_samples, max_epochs in (("a", 12, 100, 10), ("b", 8, 1, 5)):
    scale = torch.rand(n, 1, 1, 1, generator=g) * 4 + 2
    shift = torch.randn(n, 1, 1, 1, generator=g) * 2 - 5
    x = torch.randn(n, 1, 64, 96, generator=g) * scale + shift
    norm = augmentations.RunningNorm(epoch_samples=epoch_samples, max_update_epochs=max_epochs)
    y = torch.stack([norm(x[i].clone()) for i in range(n)])
    out[f"{tag}_x"] = x.numpy()
    out[f"{tag}_y"] = y.numpy()
    out[f"{tag}_cfg"] = np.array([epoch_samples, max_epochs], dtype=np.int64)
np.savez_compressed(os.path.join(HERE, "prenorm.npz"), **out)
print("wrote prenorm.npz", {k: v.shape for k, v in out.items()})
