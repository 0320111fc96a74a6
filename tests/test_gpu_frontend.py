"""GPU parity of the frontend and view kernels (through the C ABI) against the golden fixtures
generated from the reference and against the numpy oracle.  Tolerance (BASELINE.json:north_star):
log-mel within 1e-3 relative, taken as |d| <= 1e-3 * max(1, |ref|); crop and mixup indices bit-exact."""
import os
import random
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402

TOL = 1e-3
AS_STATS = (-0.8294, 4.6230)


def _args(**kw):
    base = dict(mixup=True, Gnoise=False, RRC=True, RLF=True, n_mels=64, crop_frames=96, virtual_crop_scale=[1.0, 1.5],
                local_crops_number=0, local_crops_size=[16, 16], sample_rate=16000, n_fft=1024, win_length=1024,
                hop_length=160, f_min=60, f_max=7800, unit_sec=0.95)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _relerr(a, b):
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


def test_logmel_matches_torchaudio_golden(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    mel = S.LogMelSpectrogram(16000, 1024, 1024, 160, 64, 60, 7800)
    out = mel(torch.from_numpy(g["wav"]).cuda()).cpu().numpy()
    assert out.shape == g["lms"].shape
    assert _relerr(out, g["lms"]) < TOL, _relerr(out, g["lms"])
    # mel POWER relative error where the signal is not at the eps floor
    loud = g["lms"] > -10
    assert np.abs(np.expm1(out[loud] - g["lms"][loud])).max() < TOL


def test_logmel_long_clip_and_short_window(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    wav = O.synth_wave(1, 160000, seed=int(g["wav10_seed"]))
    mel = S.LogMelSpectrogram(16000, 1024, 1024, 160, 64, 60, 7800)
    out = mel(torch.from_numpy(wav).cuda()).cpu().numpy()
    assert tuple(out.shape) == tuple(g["lms10_shape"])
    assert _relerr(out[0][:, g["lms10_frames"]], g["lms10_cols"]) < TOL
    assert _relerr(out, O.log_mel(wav)) < TOL                     # every frame, against the oracle
    mel400 = S.LogMelSpectrogram(16000, 1024, 400, 160, 64, 60, 7800)
    out400 = mel400(torch.from_numpy(g["wav"][:1]).cuda()).cpu().numpy()
    assert _relerr(out400, g["lms_win400"]) < TOL


@pytest.mark.parametrize("n_mels", [40, 80, 128])
def test_logmel_other_band_counts(golden_dir, n_mels):
    """n_mels other than the default 64 (a flag of the reference's hyperparameters.py) against torchaudio's output."""
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "logmel_nmels.npz"))
    _, f_min, f_max = (int(v) for v in g[f"cfg_{n_mels}"])
    mel = S.LogMelSpectrogram(16000, 1024, 1024, 160, n_mels, f_min, f_max)
    out = mel(torch.from_numpy(g["wav"]).cuda()).cpu().numpy()
    ref = g[f"lms_{n_mels}"]
    assert out.shape == ref.shape
    assert _relerr(out, ref) < TOL, _relerr(out, ref)
    # the z-scored crop-first path writes the same values into caller-provided slots
    cfg = O.MelConfig(n_mels=n_mels, f_min=float(f_min), f_max=float(f_max))
    meln = S.LogMelSpectrogram(16000, 1024, 1024, 160, n_mels, f_min, f_max, norm_stats=AS_STATS)
    outn = meln(torch.from_numpy(g["wav"]).cuda()).cpu().numpy()
    assert _relerr(outn, O.normalise(O.log_mel(g["wav"], cfg), AS_STATS)) < TOL


def test_logmel_ragged_batch_shapes_and_errors():
    import ssl_audio_b200 as S
    mel = S.LogMelSpectrogram(16000, 1024, 1024, 160, 64, 60, 7800, norm_stats=AS_STATS)
    for L in (513, 1000, 15200, 16001):
        wav = O.synth_wave(2, L, seed=L)
        out = mel(torch.from_numpy(wav).cuda()).cpu().numpy()
        ref = O.normalise(O.log_mel(wav), AS_STATS)
        assert out.shape == ref.shape == (2, 64, 1 + L // 160)
        assert _relerr(out, ref) < TOL, (L, _relerr(out, ref))
    assert mel(torch.zeros(0, 2000).cuda()).shape == (0, 64, 13)
    with pytest.raises(ValueError):
        mel(torch.zeros(1, 512).cuda())                            # reflect pad needs L > n_fft/2, like torch.stft
    with pytest.raises(RuntimeError):
        mel(torch.zeros(1, 2000))                                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        S.LogMelSpectrogram(16000, 512, 512, 160, 64, 60, 7800)(torch.zeros(1, 2000).cuda())


def test_pair_transform_sequence_matches_reference(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "views.npz"))
    seed = int(g["seed"])
    x = torch.from_numpy(g["x"]).cuda()
    # (a) whole batch in one call
    np.random.seed(seed); random.seed(seed)
    tf = S.AudioPairTransform(_args())
    v = tf(x)
    assert len(v) == 2 and tuple(v[0].shape) == (6, 1, 64, 96)
    got = torch.stack(v, 1).cpu().numpy()
    assert np.abs(got - g["views"]).max() < TOL, np.abs(got - g["views"]).max()
    assert tf.memory_bank_len == 12
    # (b) sample by sample, like Dataset.__getitem__ does
    np.random.seed(seed); random.seed(seed)
    tf = S.AudioPairTransform(_args())
    for b in range(6):
        crops = tf(x[b])
        assert tuple(crops[0].shape) == (1, 64, 96)
        assert np.abs(torch.stack(crops).cpu().numpy() - g["views"][b]).max() < TOL
    # the global generators advanced exactly as in the reference run
    a, c = np.random.random(), random.random()
    np.random.seed(seed); random.seed(seed)
    st = O.MixupState()
    for b in range(6):
        O.audio_pair_transform(g["x"][b], O.PairTransformConfig(), st)
    assert a == np.random.random() and c == random.random()


def test_lms_path_time_crop(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "views.npz"))
    np.random.seed(77); random.seed(77)
    fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms")
    v = fe.forward_lms(torch.from_numpy(g["lms_full"]).cuda())
    got = torch.stack(v, 1).cpu().numpy()
    assert np.abs(got - g["lms_views"]).max() < TOL


def test_multicrop_local_views(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "views.npz"))
    np.random.seed(5); random.seed(5)
    tf = S.AudioPairTransform(_args(mixup=False, local_crops_number=2))
    crops = tf(torch.from_numpy(g["x"][0]).cuda())
    assert [tuple(c.shape) for c in crops] == [(1, 64, 96)] * 2 + [(1, 16, 16)] * 2
    assert np.abs(torch.stack(crops[:2]).cpu().numpy() - g["mc_global"]).max() < TOL
    assert np.abs(torch.stack(crops[2:]).cpu().numpy() - g["mc_local"]).max() < TOL


def test_standalone_modules_match_oracle(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "views.npz"))
    x = g["x"][:3]
    xd = torch.from_numpy(x).cuda()
    np.random.seed(3); random.seed(3)
    rrc = S.RandomResizeCrop()
    got = rrc(xd).cpu().numpy()
    np.random.seed(3); random.seed(3)
    for b in range(3):
        ref, _ = O.random_resize_crop(x[b])
        assert np.abs(got[b] - ref).max() < TOL
    np.random.seed(4)
    rlf = S.RandomLinearFader()
    got = rlf(xd[0]).cpu().numpy()
    np.random.seed(4)
    ref, _ = O.linear_fader(x[0])
    assert np.abs(got - ref).max() < 1e-6
    np.random.seed(8)
    mix = S.MixupBYOLA()
    outs = [mix(xd[b]).cpu().numpy() for b in range(3)]
    np.random.seed(8)
    st = O.MixupState()
    for b in range(3):
        ref, _ = O.mixup_byola(x[b], st)
        assert np.abs(outs[b] - ref).max() < 1e-4
    assert len(mix.memory_bank) == 3 and torch.equal(mix.memory_bank[0], xd[0])
    lm = S.log_mixup_exp(xd[0], xd[1], 0.7).cpu().numpy()
    assert np.abs(lm - O.log_mixup_exp(x[0], x[1], 0.7)).max() < 1e-4
    # identity crop is an exact copy (reference property)
    i, j, h, w = S.RandomResizeCrop.get_params((64, 144), (64, 96), (0.6, 1.5), (0.6, 1.5))
    assert 1 <= h <= 64 and 1 <= w <= 144 and 0 <= i <= 64 - h and 0 <= j <= 144 - w


def test_crop_first_equals_full_then_crop_and_wav_path():
    import ssl_audio_b200 as S
    wav = O.synth_wave(5, 48000, seed=11)
    wd = torch.from_numpy(wav).cuda()
    outs = {}
    for mode in ("crop", "full"):
        np.random.seed(21); random.seed(21)
        fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode=mode)
        outs[mode] = torch.stack(fe(wd), 1).cpu().numpy()
    assert np.abs(outs["crop"] - outs["full"]).max() < 1e-4
    # against the oracle's per-sample replay of AudioSet.__getitem__ on the full log-mel
    np.random.seed(21); random.seed(21)
    st = O.MixupState()
    lms = O.log_mel(wav)
    for b in range(5):
        views, _, _ = O.frontend_clip_lms_path(lms[b], AS_STATS, O.PairTransformConfig(), st)
        assert np.abs(outs["crop"][b] - np.stack(views)).max() < 2e-3
    # wav path (datasets.py:98-122): short clip is centre padded, long clip is unit-cropped
    for L in (9000, 20000):
        w = O.synth_wave(3, L, seed=L)
        np.random.seed(2); random.seed(2)
        fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="wav")
        got = torch.stack(fe(torch.from_numpy(w).cuda()), 1).cpu().numpy()
        np.random.seed(2); random.seed(2)
        st = O.MixupState()
        for b in range(3):
            views, _, _ = O.frontend_clip_wav_path(w[b], O.MelConfig(), 0.95, AS_STATS, O.PairTransformConfig(), st)
            assert np.abs(got[b] - np.stack(views)).max() < 2e-3, (L, b)


def test_short_clip_is_padded_before_normalisation():
    import ssl_audio_b200 as S
    wav = O.synth_wave(2, 8000, seed=1)                     # 51 frames < 96
    np.random.seed(0); random.seed(0)
    fe = S.BatchFrontend(_args(mixup=False, RRC=False, RLF=False), norm_stats=AS_STATS, path="lms")
    v = fe(torch.from_numpy(wav).cuda())
    ref, _ = O.lms_trim_pad(O.log_mel(wav)[:, None], 96)
    ref = O.normalise(ref, AS_STATS)
    assert np.abs(v[0].cpu().numpy() - ref).max() < TOL
    assert torch.equal(v[0], v[1])


def test_pinned_host_waveforms_zero_copy_span_path_is_bit_identical():
    """forward(pinned host wav) plans the crop first and reads only the cropped span over PCIe; the result must
    equal forward(device wav) bit for bit (also when crops touch either end of the clip: reflect padding)."""
    import ssl_audio_b200 as S
    wav = O.synth_wave(48, 32000, seed=4)                   # 201 frames: 105 possible crop starts, many near the edges
    outs = []
    for host in (False, True):
        np.random.seed(5); random.seed(5)
        fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="crop")
        w = torch.from_numpy(wav).pin_memory() if host else torch.from_numpy(wav).cuda()
        for _ in range(2):                                   # second call: Mixup partners come from the ring
            v = fe(w)
        outs.append((torch.stack(v, 1).cpu().numpy(), fe.last_plan.starts.copy()))
    assert outs[0][1].min() <= 2 and outs[0][1].max() >= 102, "seed no longer exercises both clip edges"
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][0], outs[1][0])
    # forced edge crops through the C ABI: first and last possible crop start
    from ssl_audio_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    mel = S.LogMelSpectrogram(16000, 1024, 1024, 160, 64, 60, 7800, norm_stats=AS_STATS)
    dev = torch.device("cuda", 0)
    plan = mel.plan(dev)
    wh = torch.from_numpy(wav[:4]).pin_memory()
    wd = wh.cuda()
    starts = torch.tensor([0, 105, 1, 104], dtype=torch.int32, device=dev)
    ref = torch.empty((4, 64 * 96), device=dev)
    got = torch.empty((4, 64 * 96), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.abt_logmel_crop_fwd(plan, wd.data_ptr(), 32000, None, 4, 32000, starts.data_ptr(), 96, ref.data_ptr(), None, 64 * 96, st))
    n = C.c_int()
    _lib.check(lib.abt_wav_span_len(plan, 96, C.byref(n)))
    spans = torch.empty((4, n.value), device=dev)
    origin = torch.empty((4,), dtype=torch.int32, device=dev)
    _lib.check(lib.abt_wav_span_gather(plan, wh.data_ptr(), 1, 32000, 4, 32000, starts.data_ptr(), 96, spans.data_ptr(), origin.data_ptr(), st))
    _lib.check(lib.abt_logmel_span_fwd(plan, spans.data_ptr(), origin.data_ptr(), 4, 32000, starts.data_ptr(), 96, got.data_ptr(), None, 64 * 96, st))
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
    assert origin.cpu().tolist()[0] == 0 and origin.cpu().tolist()[1] == 32000 - n.value
    with pytest.raises(RuntimeError):
        S.BatchFrontend(_args(), norm_stats=AS_STATS)(torch.from_numpy(wav))          # pageable host memory: refused


def test_host_prefetch_pipeline_equals_sequential_calls():
    """prepare(next host batch) on a side stream while the current batch is launched (the e2e loop of bench.py) must give exactly
    what sequential forward() calls give: same draws, same ring contents, double-buffered spans never overwritten early."""
    import ssl_audio_b200 as S
    batches = [torch.from_numpy(O.synth_wave(32, 32000, seed=20 + k)).pin_memory() for k in range(5)]
    np.random.seed(9); random.seed(9)
    fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="crop")
    seq = [torch.stack(fe(w), 1).clone() for w in batches]
    torch.cuda.synchronize()
    np.random.seed(9); random.seed(9)
    fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="crop")
    side, main = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(side):
        handle = fe.prepare(batches[0])
    pipe = []
    for k in range(len(batches)):
        with torch.cuda.stream(main):
            v = fe.launch(handle)
            pipe.append(torch.stack(v, 1).clone())
        if k + 1 < len(batches):
            with torch.cuda.stream(side):
                handle = fe.prepare(batches[k + 1])
    torch.cuda.synchronize()
    for a, b in zip(seq, pipe):
        assert torch.equal(a, b)


def test_stale_prepared_batch_is_refused():
    """A prepared batch aliases a slot of the planner's 4-deep staging ring: launching it after the slot has been handed out again
    must raise instead of silently using another batch's parameters."""
    import ssl_audio_b200 as S
    fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="crop")
    wav = torch.from_numpy(O.synth_wave(8, 32000, seed=3)).cuda()
    first = fe.prepare(wav)
    later = [fe.prepare(wav) for _ in range(5)]
    with pytest.raises(RuntimeError, match="stale batch plan"):
        fe.launch(first)
    fe.launch(later[-1])                                  # recent handles are still live
    torch.cuda.synchronize()


def test_stage_pipelines_run_like_the_reference_attributes():
    """tfm.global_transform(x) / tfm.local_transform(x) (utils/transforms.py:15-46) on one log-mel: shapes, and the chain equals
    the standalone stages applied one after the other with the same draws."""
    import ssl_audio_b200 as S
    x = torch.from_numpy(O.normalise(O.log_mel(O.synth_wave(1, 16000, seed=8))[:, None][..., :96], AS_STATS)[0]).cuda().float()
    tfm = S.AudioPairTransform(_args(mixup=False))
    np.random.seed(3); random.seed(3)
    g = tfm.global_transform(x)
    l = tfm.local_transform(x)
    assert tuple(g.shape) == (1, 64, 96) and tuple(l.shape) == (1, 16, 16)
    np.random.seed(3); random.seed(3)
    ref = S.RandomLinearFader()(S.RandomResizeCrop((64, 96), virtual_crop_scale=(1.0, 1.5), freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5))(x))
    assert torch.equal(g, ref)


def test_bench_size_frontend_1024_clips_of_10s():
    """BASELINE config 2 at full size (1024 clips x 10 s, crop-first): crop starts equal the reference's np.random.randint draws
    interleaved with the view draws (replayed by the oracle), eight spot-checked clips match the oracle's per-sample path, every
    output is finite, and the full-log-mel mode agrees with crop-first on the same seed."""
    import ssl_audio_b200 as S
    B, L = 1024, 160000
    g = torch.Generator(device="cuda").manual_seed(0)
    wav = (0.1 * torch.randn(B, L, device="cuda", generator=g)).clamp_(-1, 1)
    t = torch.arange(L, device="cuda", dtype=torch.float32) / 16000.0
    wav += 0.3 * torch.sin(6.2831853 * (100.0 + 6900.0 * torch.rand(B, 1, device="cuda", generator=g)) * t[None, :])
    np.random.seed(3); random.seed(3)
    fe = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="crop")
    views = fe(wav)
    starts = fe.last_plan.starts.copy()
    v = torch.stack(views, 1)
    assert bool(torch.isfinite(v).all())
    assert starts.min() >= 0 and starts.max() <= 1001 - 97
    # oracle replay of the first 8 clips (the draws are sequential, so a prefix is enough): same crop starts, same views
    sub = wav[:8].cpu().numpy()
    np.random.seed(3); random.seed(3)
    st = O.MixupState()
    lms = O.log_mel(sub)
    for b in range(8):
        ref, rec, _ = O.frontend_clip_lms_path(lms[b], AS_STATS, O.PairTransformConfig(), st)
        assert int(starts[b]) == int(rec["start"])
        assert np.abs(v[b].cpu().numpy() - np.stack(ref)).max() < 2e-3, b
    # crop-first == full log-mel then crop (same seed -> same draws)
    np.random.seed(3); random.seed(3)
    fe2 = S.BatchFrontend(_args(), norm_stats=AS_STATS, path="lms", mode="full")
    v2 = torch.stack(fe2(wav), 1)
    assert np.array_equal(fe2.last_plan.starts, starts)
    assert float((v - v2).abs().max()) < 1e-3


def test_normalize_batch_matches_reference_golden(golden_dir):
    """NormalizeBatch (--post_norm, main.py:62-66) against outputs of the reference module: a batch of log-mels, a multi-channel
    batch with a large offset (cancellation), and a constant batch (std clamped to eps)."""
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "postnorm.npz"))
    mod = S.NormalizeBatch()
    for tag in ("a", "b"):
        y = mod(torch.from_numpy(g[f"{tag}_x"]).cuda()).cpu().numpy()
        ref = g[f"{tag}_y"]
        assert np.abs(y - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), tag
        assert np.abs(y - O.normalize_batch(g[f"{tag}_x"])).max() <= 1e-3 * max(1.0, np.abs(ref).max())
    y = mod(torch.from_numpy(g["c_x"]).cuda()).cpu().numpy()       # constant input: (x - mean) / eps with mean == x -> exactly 0
    assert np.array_equal(y, g["c_y"])
    # bench-size batch: statistics of the result are (0, 1)
    x = torch.randn(1024, 1, 64, 96, device="cuda") * 4.6 - 0.8
    y = mod(x)
    assert abs(float(y.mean())) < 1e-4 and abs(float(y.std()) - 1.0) < 1e-4
    with pytest.raises(RuntimeError):
        mod(torch.zeros(2, 1, 64, 96))


def test_running_norm_matches_reference_golden(golden_dir):
    """RunningNorm (--pre_norm, main.py:272-277) against the reference module fed sample by sample: a whole batch in one call, the same
    sequence split over several calls (the running state is carried on the device), single (1, F, T) samples, and the frozen
    statistics after max_update samples."""
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "prenorm.npz"))
    for tag in ("a", "b"):
        epoch_samples, max_epochs = (int(v) for v in g[f"{tag}_cfg"])
        x, ref = torch.from_numpy(g[f"{tag}_x"]).cuda(), g[f"{tag}_y"]
        y = S.RunningNorm(epoch_samples, max_epochs)(x).cpu().numpy()
        np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5, err_msg=tag)
        np.testing.assert_allclose(y, O.running_norm(g[f"{tag}_x"], epoch_samples, max_epochs), rtol=0, atol=2e-5)
        mod = S.RunningNorm(epoch_samples, max_epochs)
        parts = [mod(x[:3]), mod(x[3]).unsqueeze(0), mod(x[4:])]
        np.testing.assert_allclose(torch.cat(parts).cpu().numpy(), ref, rtol=0, atol=2e-5, err_msg=tag + " split")
    with pytest.raises(RuntimeError):
        S.RunningNorm(10)(torch.zeros(2, 1, 64, 96))


def test_mix_gaussian_noise_matches_reference_golden(golden_dir):
    """MixGaussianNoise (args.Gnoise): the bare module and the whole Mixup -> Gnoise -> RRC -> RLF chain against the reference, fed the
    reference's own N(0, 1) draws; then the default path, whose draws come from torch's CUDA generator, against the oracle fed the
    same draws."""
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "gnoise.npz"))
    np.random.seed(41)
    mod = S.MixGaussianNoise(ratio=0.2)
    y = mod(torch.from_numpy(g["m_x"]).cuda(), torch.from_numpy(g["m_n"]).cuda()).cpu().numpy()
    assert np.abs(y - g["m_y"]).max() < TOL
    seed = int(g["seed"])
    x, noise = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["noise"]).cuda()
    np.random.seed(seed); random.seed(seed)
    v = S.AudioPairTransform(_args(Gnoise=True))(x, noise)
    got = torch.stack(v, 1).cpu().numpy()
    assert np.abs(got - g["views"]).max() < TOL, np.abs(got - g["views"]).max()
    np.random.seed(seed); random.seed(seed)
    tf = S.AudioPairTransform(_args(Gnoise=True))
    for b in range(6):                                         # sample by sample, like Dataset.__getitem__
        crops = tf(x[b], noise[b:b + 1])
        assert np.abs(torch.stack(crops).cpu().numpy() - g["views"][b]).max() < TOL
    # default: noise drawn on the device
    np.random.seed(7); torch.manual_seed(7)
    y = S.MixGaussianNoise(0.2)(x[0])
    torch.manual_seed(7)
    n = torch.randn(1, 1, 64, 96, device="cuda").cpu().numpy()[0]
    np.random.seed(7)
    ref, _ = O.mix_gaussian_noise(g["x"][0], 0.2, n)
    assert tuple(y.shape) == (1, 64, 96) and np.abs(y.cpu().numpy() - ref).max() < TOL
    with pytest.raises(ValueError):
        S.AudioPairTransform(_args(Gnoise=True))(x, noise[:3])
