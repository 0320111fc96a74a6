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

__all__ = [
    "AudioPairTransform", "RandomResizeCrop", "RandomLinearFader", "MixupBYOLA", "MixGaussianNoise", "NormalizeBatch", "RunningNorm", "log_mixup_exp",
    "LogMelSpectrogram", "BatchFrontend", "BarlowTwinsLoss", "bt_loss_fwd_bwd", "off_diagonal",
    "LARS", "EMA", "update_moving_average",
]
