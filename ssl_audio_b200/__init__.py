"""ssl_audio_b200 -- B200-native (sm_100a) hot path of Audio Barlow Twins.

Host-side mirror of the reference's interfaces for the path named in BASELINE.json:
    AudioPairTransform, RandomResizeCrop, RandomLinearFader, MixupBYOLA, log_mixup_exp   (views)
    LogMelSpectrogram, BatchFrontend                                                     (frontend)
    BarlowTwinsLoss, off_diagonal                                                        (objective)
    LARS, EMA, update_moving_average                                                     (step-adjacent optimiser math)
All compute runs in hand-written CUDA behind the C ABI of include/abt_b200.h; importing this package
never falls back to PyTorch/CPU implementations.
"""
from .augmentations import MixGaussianNoise, MixupBYOLA, NormalizeBatch, RandomLinearFader, RandomResizeCrop, RunningNorm, log_mixup_exp
from .frontend import BatchFrontend, LogMelSpectrogram
from .loss import BarlowTwinsLoss, bt_loss_fwd_bwd, off_diagonal
from .optim import EMA, LARS, update_moving_average
from .transforms import AudioPairTransform


def set_reserved_sms(n_sms: int) -> int:
    """SMs the persistent tensor-core kernels of the objective leave free (default 0) for work that must run beside them -- the
    pinned-host span gather of a prefetching input pipeline (`BatchFrontend.prepare(host_wav)` on a side stream occupies 4 SMs).
    Returns the previous value.  include/abt_b200.h: abt_set_reserved_sms."""
    from . import _lib
    return int(_lib.load().abt_set_reserved_sms(int(n_sms)))

__all__ = [
    "AudioPairTransform", "RandomResizeCrop", "RandomLinearFader", "MixupBYOLA", "MixGaussianNoise", "NormalizeBatch", "RunningNorm", "log_mixup_exp",
    "LogMelSpectrogram", "BatchFrontend", "BarlowTwinsLoss", "bt_loss_fwd_bwd", "off_diagonal",
    "LARS", "EMA", "update_moving_average", "set_reserved_sms",
]
