"""`AudioPairTransform` exposes the reference's stage pipelines (utils/transforms.py:15-46) as attributes."""
import types

import torch.nn as nn

import ssl_audio_b200 as S


def _args(**kw):
    base = dict(mixup=True, Gnoise=True, RRC=True, RLF=True, n_mels=64, crop_frames=96, virtual_crop_scale=[1.0, 1.5],
                local_crops_number=2, local_crops_size=[16, 16])
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_stage_pipelines_mirror_the_reference_structure():
    t = S.AudioPairTransform(_args())
    assert isinstance(t.global_transform, nn.Sequential) and isinstance(t.local_transform, nn.Sequential)
    assert [type(m).__name__ for m in t.global_transform] == ["MixupBYOLA", "MixGaussianNoise", "RandomResizeCrop", "RandomLinearFader"]
    assert [type(m).__name__ for m in t.local_transform] == ["RandomResizeCrop"]
    rrc = t.global_transform[2]
    assert "virtual_crop_size=(1.0, 1.5)" in repr(rrc) and "time_scale=(0.6, 1.5)" in repr(rrc)
    assert "time_scale=(0.05, 0.6)" in repr(t.local_transform[0])
    t2 = S.AudioPairTransform(_args(mixup=False, Gnoise=False, RLF=False))
    assert [type(m).__name__ for m in t2.global_transform] == ["RandomResizeCrop"]
    assert isinstance(S.AudioPairTransform(_args(), train_transform=False).global_transform, nn.Identity)
    assert list(t.state_dict().keys()) == []          # as in the reference: no parameters, no buffers
