"""Host planner (csrc/planner.cpp): bit-exact replay of numpy's legacy global generator and CPython's
`random` in the reference's draw order, checked against the oracle (which calls np.random / random
directly) and against the golden fixtures recorded from the reference."""
import os
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import abt_oracle as O
from ssl_audio_b200.planner import ViewPlanner


def _boxes(params):
    return np.stack([params[k] for k in ("i", "j", "h", "w")], -1).reshape(-1, 4)


def test_golden_sequence_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "views.npz"))
    seed = int(g["seed"])
    np.random.seed(seed); random.seed(seed)
    pl = ViewPlanner(mixup=True, rrc=True, rlf=True)
    p = pl.plan(6).params
    np.testing.assert_array_equal(_boxes(p), g["boxes"])
    flat = p.reshape(-1)
    alphas = 0.2 * g["mix_u"]
    np.testing.assert_array_equal(flat["w_x"][1:], (1.0 - alphas).astype(np.float32)[1:])
    np.testing.assert_array_equal(flat["w_z"][1:], (1.0 - (1.0 - alphas)).astype(np.float32)[1:])
    assert flat["z_kind"][0] == 0 and flat["w_x"][0] == 1.0          # empty bank: mixed = x
    fades = (2.0 * g["fade_u"] - 1.0).astype(np.float32)
    np.testing.assert_array_equal(np.stack([flat["head"], flat["tail"]], -1), fades)
    # partner uid = memory_bank[idx] where the bank holds each clip twice
    hist = []
    for k in range(12):
        if hist:
            assert flat["z_kind"][k] == 2 and flat["z_index"][k] == hist[int(g["bank_idx"][k - 1])]
        hist.append(k // 2)
    assert pl.bank_len() == 12


def test_global_generators_advance_like_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "views.npz"))
    np.random.seed(9); random.seed(9)
    ViewPlanner(mixup=True, rrc=True, rlf=True).plan(4, time_crop_range=205)
    after = (np.random.random(), np.random.randint(1000), random.random(), random.randint(0, 99), np.random.randn())
    np.random.seed(9); random.seed(9)
    stt = O.MixupState()
    for b in range(4):
        O.frontend_clip_lms_path(np.zeros((64, 301), np.float32), None, O.PairTransformConfig(), stt)
    ref = (np.random.random(), np.random.randint(1000), random.random(), random.randint(0, 99), np.random.randn())
    assert after == ref


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n_clips=st.integers(1, 40), n_memory=st.integers(1, 9),
       crop=st.integers(0, 300), wavcrop=st.integers(0, 5000))
def test_random_configs_match_oracle_replay(seed, n_clips, n_memory, crop, wavcrop):
    np.random.seed(seed); random.seed(seed)
    pl = ViewPlanner(mixup=True, rrc=True, rlf=True, n_memory=n_memory, ring_slots=n_memory + 64, n_local=1)
    bp = pl.plan(n_clips, time_crop_range=crop, wav_crop_range=wavcrop)
    np.random.seed(seed); random.seed(seed)
    bank = []          # uids
    for b in range(n_clips):
        if wavcrop > 0:
            assert bp.wav_starts[b] == random.randint(0, wavcrop)
        else:
            assert bp.wav_starts[b] == -1
        if crop > 0:
            assert bp.starts[b] == np.random.randint(crop)
        else:
            assert bp.starts[b] == -1
        for v in range(2):
            pv = bp.params[b, v]
            alpha = 0.2 * np.random.random()
            if bank:
                idx = np.random.randint(len(bank))
                uid = bank[idx]
                assert (pv["z_kind"], pv["z_index"]) == (2, uid)     # every partner is in this batch here
                assert pv["w_x"] == np.float32(1.0 - alpha) and pv["w_z"] == np.float32(1.0 - (1.0 - alpha))
            else:
                assert pv["z_kind"] == 0
            bank = (bank + [b])[-n_memory:]
            i, j, h, w = O.rrc_get_params((64, 144), (64, 96), (0.6, 1.5), (0.6, 1.5))
            assert (pv["i"], pv["j"], pv["h"], pv["w"]) == (i, j, h, w)
            head, tail = 2.0 * np.random.rand(2) - 1.0
            assert pv["head"] == np.float32(head) and pv["tail"] == np.float32(tail)
            assert pv["flags"] == 7
        i, j, h, w = O.rrc_get_params((64, 96), (64, 96), (0.05, 0.6), (0.05, 0.6))
        pv = bp.params[b, 2]
        assert (pv["i"], pv["j"], pv["h"], pv["w"], pv["flags"]) == (i, j, h, w, 2)
    assert pl.bank_len() == min(n_memory, 2 * n_clips)


def test_ring_slots_across_batches():
    np.random.seed(0); random.seed(0)
    pl = ViewPlanner(mixup=True, rrc=False, rlf=False, n_memory=8, ring_slots=8 + 4)
    seen_slots = {}
    for step in range(6):
        bp = pl.plan(4)
        first = step * 4
        for b in range(4):
            assert bp.slots[b] == (first + b) % 12
            for v in range(2):
                pv = bp.params[b, v]
                if pv["z_kind"] == 1:      # an older clip: its slot must still hold that clip
                    uid = seen_slots[int(pv["z_index"])]
                    assert first - 8 <= uid < first
                elif pv["z_kind"] == 2:
                    assert 0 <= pv["z_index"] <= b
        for b in range(4):
            seen_slots[int(bp.slots[b])] = first + b
    with pytest.raises(ValueError):
        pl.plan(5)                          # batch larger than ring_slots - n_memory


def test_rrc_ranges_property():
    np.random.seed(1); random.seed(1)
    p = ViewPlanner(mixup=False, rrc=True, rlf=False).plan(2000).params.reshape(-1)
    assert p["h"].min() >= 38 and p["h"].max() <= 64
    assert p["w"].min() >= 57 and p["w"].max() <= 143
    assert (p["i"] >= 0).all() and (p["i"] + p["h"] <= 64).all()
    assert (p["j"] >= 0).all() and (p["j"] + p["w"] <= 144).all()
    assert 0.45 < (p["h"] == 64).mean() < 0.65          # P(h = 64) ~ 55.8 %


def test_gaussian_noise_draw_order(golden_dir):
    """args.Gnoise adds one np.random.rand() per global view between the Mixup and the RandomResizeCrop draws
    (augmentations.py:136, utils/transforms.py:18-34): lambd and everything drawn after it must match the oracle replay."""
    g = np.load(os.path.join(golden_dir, "gnoise.npz"))
    seed = int(g["seed"])
    np.random.seed(seed); random.seed(seed)
    pl = ViewPlanner(mixup=True, rrc=True, rlf=True, gnoise=True)
    p = pl.plan(6).params
    after = (np.random.random(), random.random())
    np.random.seed(seed); random.seed(seed)
    st_, cfg = O.MixupState(), O.PairTransformConfig(Gnoise=True)
    for b in range(6):
        _, recs = O.audio_pair_transform(g["x"][b], cfg, st_, g["noise"][b][:, None])
        for v, r in enumerate(recs):
            pv = p[b, v]
            assert pv["flags"] == 15
            assert pv["g_lambda"] == np.float32(r["lambd"]) and pv["g_keep"] == np.float32(1 - r["lambd"])
            assert (pv["i"], pv["j"], pv["h"], pv["w"]) == (r["i"], r["j"], r["h"], r["w"])
            assert pv["head"] == np.float32(r["head"]) and pv["tail"] == np.float32(r["tail"])
    assert after == (np.random.random(), random.random())
    # without the flag nothing changes: no draw, no flag
    np.random.seed(3); random.seed(3)
    q = ViewPlanner(mixup=True, rrc=True, rlf=True).plan(2).params
    assert (q["flags"] == 7).all() and (q["g_lambda"] == 0).all()
