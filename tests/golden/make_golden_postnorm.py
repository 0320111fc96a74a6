"""Golden fixture for NormalizeBatch FROM THE LIVE REFERENCE (build container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_postnorm.py
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
g = torch.Generator().manual_seed(11)
out = {}
for tag, shape, scale, shift in (("a", (5, 1, 64, 96), 4.6, -0.8), ("b", (3, 3, 16, 16), 0.01, 100.0), ("c", (2, 1, 64, 96), 0.0, 1.5)):
    x = torch.randn(shape, generator=g) * scale + shift
    y = augmentations.NormalizeBatch()(x)
    out[f"{tag}_x"] = x.numpy()
    out[f"{tag}_y"] = y.numpy()
np.savez_compressed(os.path.join(HERE, "postnorm.npz"), **out)
print("wrote postnorm.npz", {k: v.shape for k, v in out.items()})
