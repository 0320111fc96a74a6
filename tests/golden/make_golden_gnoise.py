"""Golden fixture for MixGaussianNoise / args.Gnoise FROM THE LIVE REFERENCE (build container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_gnoise.py

The reference draws its noise with torch.normal(0, lambd, shape) from torch's CPU generator.  ATen fills N(0, 1) and scales by
float32(lambd); to pin the draws, `torch.normal` inside the reference's augmentations module is replaced by a recorder that takes
torch.randn(shape) * std from the same generator -- and the script first ASSERTS that this is bit-identical to the real call.
"""
import os
import random
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import augmentations  # noqa: E402  (reference)
import utils.transforms as ref_transforms  # noqa: E402  (reference)

torch.set_num_threads(1)

# torch.normal(0, s, shape) == torch.randn(shape) * float32(s), draw for draw
for s in (0.2, 0.013, 0.1999):
    torch.manual_seed(7)
    a = torch.normal(0, s, (1, 64, 96))
    torch.manual_seed(7)
    b = torch.randn(1, 64, 96) * s
    assert torch.equal(a, b), s

draws = []


class _TorchProxy:
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def normal(mean, std, shape):
        n = torch.randn(shape)
        draws.append(n.numpy().copy())
        return n * std + mean


augmentations.torch = _TorchProxy()
out = {}
try:
    # bare module, three single samples
    np.random.seed(41); random.seed(41); torch.manual_seed(41)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(3, 1, 64, 96, generator=g) * 1.3 - 0.2
    mod = augmentations.MixGaussianNoise(ratio=0.2)
    ys = [mod(x[b]) for b in range(3)]
    out["m_x"], out["m_y"], out["m_n"] = x.numpy(), torch.stack(ys).numpy(), np.stack(draws)
    draws.clear()

    # the whole global chain, Mixup -> Gnoise -> RRC -> RLF, six samples in sequence
    seed = 4321
    np.random.seed(seed); random.seed(seed); torch.manual_seed(seed)
    args = types.SimpleNamespace(mixup=True, Gnoise=True, RRC=True, RLF=True, n_mels=64, crop_frames=96, virtual_crop_scale=[1.0, 1.5],
                                 local_crops_number=0, local_crops_size=[16, 16])
    B = 6
    x = torch.randn(B, 1, 64, 96, generator=g) * 1.1 + 0.1
    tfm = ref_transforms.AudioPairTransform(args)
    views = [torch.stack(tfm(x[b])).numpy() for b in range(B)]
    out["seed"] = np.array(seed)
    out["x"] = x.numpy()
    out["views"] = np.stack(views)                                        # (B, 2, 1, 64, 96)
    out["noise"] = np.stack(draws).reshape(B, 2, 64, 96)                  # N(0, 1) draws, sample-major, view-minor
finally:
    augmentations.torch = torch
np.savez_compressed(os.path.join(HERE, "gnoise.npz"), **out)
print("wrote gnoise.npz", {k: v.shape for k, v in out.items()})
