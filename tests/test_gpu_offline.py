"""Offline / evaluation-time consumers of the log-mel kernel (ssl_audio_b200/offline.py) against the library calls the reference
makes for them (torchaudio MelSpectrogram, torch mean / std, np.save): old/data_manager/wav_to_lms.py:41-87, datasets.py:87-96,
118-119, 362-376, main.py:240-252, hear/sample/vit.py:90-106."""
import json
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402


def _cfg(win_length=1024):
    return types.SimpleNamespace(sample_rate=16000, n_fft=1024, win_length=win_length, hop_length=160, n_mels=64, f_min=60, f_max=7800)


def _close(got, ref, tol=1e-3):
    assert np.all(np.abs(got - ref) <= tol * np.maximum(1.0, np.abs(ref))), float(np.abs(got - ref).max())


def test_npy_cache_writer_contract(tmp_path):
    from ssl_audio_b200.offline import LogMelCacheWriter
    wav = O.synth_wave(3, 16000, seed=4)
    short = O.synth_wave(1, 9000, seed=5)[0]
    names = ["a/x.wav", "a/y.wav", "b/z.wav", "b/short.wav"]
    waves = [wav[0], wav[1], wav[2], short]
    w = LogMelCacheWriter(str(tmp_path))
    written = w.convert(names, waves)
    assert written == ["x.npy", "y.npy", "z.npy", "short.npy"]
    for nme, wv in zip(names, waves):
        arr = np.load(tmp_path / (nme[:-4] + ".npy"))
        assert arr.dtype == np.float32 and arr.shape == (64, 1 + len(wv) // 160)
        _close(arr, O.log_mel(wv[None])[0])
    stamp = os.path.getmtime(tmp_path / "a" / "x.npy")
    assert w.convert(names[:2], waves[:2]) == ["", ""]                  # "already exist": left alone
    assert os.path.getmtime(tmp_path / "a" / "x.npy") == stamp


def test_mean_std_and_calculate_norm_stats(tmp_path):
    from ssl_audio_b200.offline import calculate_norm_stats, mean_std
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 1, 64, 96, generator=g) * 4.6 - 0.83
    m, s = mean_std(x.cuda())
    assert abs(float(m) - float(x.double().mean())) < 1e-6 and abs(float(s) - float(x.double().std())) < 1e-6
    big = (torch.randn(5, 7, generator=g) * 1e-3 + 250.0)                  # |mean| >> std
    m, s = mean_std(big.cuda())
    assert abs(float(m) - float(big.double().mean())) < 1e-4 and abs(float(s) / float(big.double().std()) - 1.0) < 1e-3

    class DS:
        def __len__(self):
            return 37

        def __getitem__(self, i):
            return x[i], 0
    np.random.seed(7)
    idxs = np.random.randint(0, 37, size=300)
    stack = torch.stack([x[i] for i in idxs])
    ref = (float(stack.mean()), float(stack.std() + torch.finfo().eps))     # datasets.py:369-370
    np.random.seed(7)
    path = str(tmp_path / "norm_stats.json")
    got = calculate_norm_stats(DS(), n_norm_calc=300, json_path=path, chunk=128)
    assert abs(got[0] - ref[0]) < 1e-5 and abs(got[1] - ref[1]) < 1e-5
    assert json.load(open(path)) == {"mean": got[0], "std": got[1]}


@pytest.mark.parametrize("seconds,from_wav", [(10.0, True), (10.0, False), (3.0, True), (3.0, False)])
def test_eval_features_crop_711(seconds, from_wav):
    """main.py:240-252: crop_frames=711, transform=None -- random crop when the clip is longer (one np.random.randint per clip),
    right zero-pad BEFORE the z-score when it is shorter."""
    from ssl_audio_b200.offline import EvalFeatures
    stats = (-4.950, 5.855)
    wav = O.synth_wave(3, int(seconds * 16000), seed=8)
    lms = O.log_mel(wav)
    t_full = lms.shape[-1]
    ev = EvalFeatures(_cfg(), norm_stats=stats, crop_frames=711)
    np.random.seed(11)
    got = ev(torch.from_numpy(wav).cuda() if from_wav else torch.from_numpy(lms).cuda()).cpu().numpy()
    assert got.shape == (3, 1, 64, 711)
    np.random.seed(11)
    for b in range(3):
        ref, _ = O.lms_trim_pad(lms[b][None], 711)
        _close(got[b], O.normalise(ref, stats), 1e-3 if from_wav else 1e-6)


def test_hear_features_win_length_400():
    import torchaudio.transforms as AT
    from ssl_audio_b200.offline import HearFeatures
    cfg = _cfg(win_length=400)
    wav = O.synth_wave(4, 24000, seed=9)
    hf = HearFeatures(cfg)
    got = hf._to_normalized_spec(torch.from_numpy(wav).cuda()).cpu().numpy()
    mel = AT.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=400, hop_length=160, n_mels=64, f_min=60, f_max=7800, power=2)
    x = (mel(torch.from_numpy(wav)) + torch.finfo().eps).log().unsqueeze(1)       # hear/sample/vit.py:90-94
    ref = ((x - x.mean()) / x.std()).numpy()                                       # hear/sample/vit.py:97-100
    assert got.shape == ref.shape == (4, 1, 64, 151)
    _close(got, ref)
    ts = hf._get_timestamps(torch.from_numpy(wav), torch.zeros(4, 10, 8))
    assert ts.shape == (4, 10) and abs(float(ts[0, 1]) - 0.15) < 1e-6
