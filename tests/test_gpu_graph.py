"""The library's launches are capture-safe: a step (frontend kernels + objective forward and backward) captured once in a CUDA graph and
replayed gives what the eager calls give, with the batch's random parameters uploaded into the static plan buffer before every replay."""
import random
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402

AS_STATS = (-0.8294, 4.6230)


def _args(d):
    return types.SimpleNamespace(mixup=True, Gnoise=False, RRC=True, RLF=True, n_mels=64, crop_frames=96, virtual_crop_scale=[1.0, 1.5],
                                 local_crops_number=0, local_crops_size=[16, 16], sample_rate=16000, n_fft=1024, win_length=1024,
                                 hop_length=160, f_min=60, f_max=7800, unit_sec=0.95, projector_out_dim=d, HSIC=False, alpha=1.0, lmbda=0.005)


@pytest.mark.parametrize("n,d", [(64, 512), (256, 512)])
def test_captured_step_equals_eager_step(n, d):
    """n = 64: the one-launch objective (programmatic dependent launch inside the graph); n = 256: statistics + CORR + GRAD."""
    import ssl_audio_b200 as S
    cfg = _args(d)
    wav = torch.from_numpy(O.synth_wave(n, 32000, seed=2)).cuda()
    z1g, z2g = O.synth_embeddings(n, d, seed=3)
    z1, z2 = torch.from_numpy(z1g).cuda().bfloat16(), torch.from_numpy(z2g).cuda().bfloat16()
    steps = 4

    def fresh():
        np.random.seed(11); random.seed(11)
        return S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="crop"), S.BarlowTwinsLoss(cfg, ncrops=2).cuda()

    def eager_step(fe, crit):
        a = z1.detach().requires_grad_(True); b = z2.detach().requires_grad_(True)
        views = fe(wav)
        loss = crit(b, a, ngcrops_each=1)
        loss.backward()
        return torch.stack(views, 1).clone(), float(loss.detach()), a.grad.clone(), b.grad.clone()

    # eager: one warm-up step + steps + 1 measured ones
    fe, crit = fresh()
    eager_step(fe, crit)
    eager = [eager_step(fe, crit) for _ in range(steps + 1)]
    rm_eager = crit.bn.running_mean.clone()
    torch.cuda.synchronize()

    # captured: the same warm-up step on the same modules (lazy plans, rings, workspaces and kernel attributes must exist before the
    # capture: they allocate and copy synchronously), then the launches of one step are captured once
    fe, crit = fresh()
    eager_step(fe, crit)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    ga = z1.detach().requires_grad_(True); gb = z2.detach().requires_grad_(True)
    handle = fe.prepare(wav, static=True)            # the first measured step's plan; nothing executes during the capture itself
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cap = torch.cuda.current_stream()
        fork = torch.cuda.Event(); fork.record(cap)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            g_views = fe.launch(handle)
        g_loss = crit(gb, ga, ngcrops_each=1)
        g_loss.backward()
        cap.wait_stream(side)
    got = []
    graph.replay()                                   # the plan prepared before the capture
    got.append((torch.stack(g_views, 1).clone(), float(g_loss.detach()), ga.grad.clone(), gb.grad.clone()))
    for _ in range(steps):
        fe.prepare(wav, static=True)
        graph.replay()
        got.append((torch.stack(g_views, 1).clone(), float(g_loss.detach()), ga.grad.clone(), gb.grad.clone()))
    torch.cuda.synchronize()
    for (ev, el, e1, e2), (gv, gl, g1, g2) in zip(eager, got):
        assert torch.equal(ev, gv)                                         # same draws, same ring contents, same kernels
        assert abs(el - gl) <= 1e-6 * abs(el)
        if n <= 128:
            assert torch.equal(e1, g1) and torch.equal(e2, g2)             # the one-launch objective is deterministic
        else:
            assert (e1.float() - g1.float()).norm() <= 2e-4 * e1.float().norm()
    assert torch.allclose(crit.bn.running_mean, rm_eager, rtol=1e-5, atol=1e-6)
