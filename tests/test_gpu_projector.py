"""Projector tail fused with the objective's statistics pass (model.py:22,25-31 + utils/loss.py:15-30):
the LINEAR mode of the tensor-core kernel against the float64 oracle, through the C ABI."""
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402

TOL = 1e-3


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def _inputs(n, k, d, seed):
    rng = np.random.default_rng(seed)
    h1 = np.maximum(rng.standard_normal((n, k)), 0.0).astype(np.float32)            # post-ReLU activations
    h2 = np.maximum(0.7 * h1 + 0.7 * rng.standard_normal((n, k)), 0.0).astype(np.float32)
    w = (rng.standard_normal((d, k)) / np.sqrt(k)).astype(np.float32)
    return O.round_bf16(h1), O.round_bf16(h2), O.round_bf16(w)


def _cfg(d, hsic=False):
    return types.SimpleNamespace(projector_out_dim=d, HSIC=hsic, alpha=1.0, lmbda=0.005)


@pytest.mark.parametrize("n,k,d", [(256, 512, 512), (200, 320, 576), (1024, 1024, 2048), (130, 64, 64)])
def test_linear_outputs_and_statistics_handover(n, k, d):
    """z = round_bf16(h W^T) and the 5 column sums of the rounded outputs (ragged N, K not a multiple of the k-block, D not a
    multiple of the pair tile)."""
    from ssl_audio_b200.projector import proj_tail_fwd
    h1, h2, w = _inputs(n, k, d, seed=n + d)
    _, rz1, rz2, *_ = O.proj_tail_loss(h1, h2, w)
    t = [torch.from_numpy(x).cuda().bfloat16() for x in (h1, h2, w)]
    pack = torch.zeros(7 * d, device="cuda")
    z1, z2 = proj_tail_fwd(t[0], t[1], t[2], pack)
    torch.cuda.synchronize()
    g1, g2 = z1.float().cpu().numpy(), z2.float().cpu().numpy()
    for g, r in ((g1, rz1), (g2, rz2)):
        assert _rel(g, r) < 3e-4                       # identical up to rare one-ulp flips of the fp32 -> bf16 rounding (fp32 vs float64 sums)
        assert np.mean(g == r) > 0.98
    pk = pack.cpu().numpy().reshape(7, d).astype(np.float64)
    a, b = g1.astype(np.float64), g2.astype(np.float64)
    want = np.stack([a.sum(0), (a * a).sum(0), b.sum(0), (b * b).sum(0), (a * b).sum(0)])
    scale = np.stack([np.abs(a).sum(0), (a * a).sum(0), np.abs(b).sum(0), (b * b).sum(0), np.abs(a * b).sum(0)]) + 1e-6
    assert np.abs(pk[:5] - want).max() <= 1e-5 * scale.max()
    assert np.abs((pk[:5] - want) / scale).max() < 1e-5
    assert not pk[5:].any()                            # unshifted sums
    z1b, z2b = proj_tail_fwd(t[0], t[1], t[2], pack)   # deterministic
    assert torch.equal(z1, z1b) and torch.equal(z2, z2b)


@pytest.mark.parametrize("n,k,d,hsic", [(256, 512, 512, False), (200, 320, 576, True), (1024, 512, 1024, False)])
def test_objective_from_the_handover_matches_oracle(n, k, d, hsic):
    """Loss and d loss / dz (fp32 outputs, true 1e-3) when the objective continues from the fused statistics, never reading z for them."""
    from ssl_audio_b200 import dist as D
    from ssl_audio_b200.projector import proj_tail_fwd
    h1, h2, w = _inputs(n, k, d, seed=3 * n + d)
    t = [torch.from_numpy(x).cuda().bfloat16() for x in (h1, h2, w)]
    be = D._CUDA_BACKEND
    ws = be.workspace(t[0].device, n, 1, d, d)
    z1, z2 = proj_tail_fwd(t[0], t[1], t[2], ws["pack_all"])
    rm, rv = torch.zeros(d, device="cuda"), torch.ones(d, device="cuda")
    be.normalize(ws, z1, z2, 1, 0, d, 1e-5, 0.1, rm, rv)
    parts, dz1, dz2 = be.rows(ws, torch.float32, t[0].device, n, 1, d, 0, d, 1.0, 0.005, hsic, 1.0, 3, 0)
    torch.cuda.synchronize()
    p = parts.cpu().numpy()
    loss = 0.005 * p[0] + (2 * 0.005 * p[1] + 0.005 * d * (d - 1) if hsic else 0.0) + 1.0 * p[2]
    # oracle on the GPU's own z (isolates the objective) and on the oracle's z (end to end)
    gz1, gz2 = z1.float().cpu().numpy(), z2.float().cpu().numpy()
    rl, r1, r2, _ = O.bt_loss_forward_backward(gz1, gz2, 1.0, 0.005, hsic)
    assert abs(loss - rl) <= TOL * abs(rl), (loss, rl)
    assert _rel(dz1.cpu().numpy(), r1) < TOL and _rel(dz2.cpu().numpy(), r2) < TOL
    el, *_ = O.proj_tail_loss(h1, h2, w, 1.0, 0.005, hsic)
    assert abs(loss - el) <= TOL * abs(el), (loss, el)
    m, v = O.bn_running_update(np.zeros(d), np.ones(d), gz1)
    m, v = O.bn_running_update(m, v, gz2)
    assert np.abs(rm.cpu().numpy() - m).max() < 1e-4 and np.abs(rv.cpu().numpy() - v).max() < 1e-3


def test_autograd_node_equals_linear_plus_loss_module():
    """proj_tail_forward_loss == loss_module.forward_loss(F.linear(h_a, W), F.linear(h_b, W)): loss, dh_a, dh_b, dW; state_dict too."""
    import ssl_audio_b200 as S
    from ssl_audio_b200.projector import proj_tail_forward_loss
    n, k, d = 384, 512, 1024
    h1, h2, w = _inputs(n, k, d, seed=5)
    el, _, _, _, _, edh1, edh2, edw = O.proj_tail_loss(h1, h2, w)
    outs = []
    for fused in (True, False):
        a = torch.from_numpy(h1).cuda().bfloat16().requires_grad_(True)
        b = torch.from_numpy(h2).cuda().bfloat16().requires_grad_(True)
        W = torch.from_numpy(w).cuda().bfloat16().requires_grad_(True)
        crit = S.BarlowTwinsLoss(_cfg(d), ncrops=2).cuda()
        if fused:
            loss = proj_tail_forward_loss(crit, a, b, W)
        else:
            loss = crit.forward_loss(torch.nn.functional.linear(a, W), torch.nn.functional.linear(b, W))
        (3.0 * loss).backward()
        torch.cuda.synchronize()
        outs.append((float(loss), a.grad.float().cpu().numpy(), b.grad.float().cpu().numpy(), W.grad.float().cpu().numpy(),
                     {k_: v_.float().cpu().numpy() for k_, v_ in crit.state_dict().items()}))
    (lf, a1, b1, w1, sd1), (lu, a2, b2, w2, sd2) = outs
    assert abs(lf - el) <= TOL * abs(el) and abs(lf - lu) <= 2e-4 * abs(lu)
    # dh / dW are bf16 outputs of bf16 GEMMs over bf16 dz in BOTH paths (rounding quantum ~2e-3): they must agree with each other and
    # with the float64 oracle at that quantum; the kernel-level 1e-3 statement is test_objective_from_the_handover_matches_oracle
    for got, other, ref in ((a1, a2, 3.0 * edh1), (b1, b2, 3.0 * edh2), (w1, w2, 3.0 * edw)):
        e_fused, e_lib = _rel(got, ref), _rel(other, ref)
        assert e_fused < 6e-3 and e_fused <= 1.1 * e_lib + 5e-4, (e_fused, e_lib)      # measured 4.1e-3 against 4.6e-3 for nn.Linear + the loss module (grad_output = 3 re-rounds dz in both)
    assert sd1.keys() == sd2.keys()
    for key in sd1:
        assert np.allclose(sd1[key], sd2[key], rtol=1e-4, atol=1e-5), key


def test_repeated_calls_share_a_workspace_without_accumulating():
    """Same shape twice: the on-diagonal sum of the second evaluation must not contain the first one's."""
    import ssl_audio_b200 as S
    from ssl_audio_b200.projector import proj_tail_forward_loss
    n, k, d = 256, 128, 512
    crit = S.BarlowTwinsLoss(_cfg(d), ncrops=2).cuda()
    losses = []
    for seed in (1, 2, 1):
        h1, h2, w = _inputs(n, k, d, seed=seed)
        t = [torch.from_numpy(x).cuda().bfloat16() for x in (h1, h2, w)]
        losses.append((float(proj_tail_forward_loss(crit, t[0], t[1], t[2])), O.proj_tail_loss(h1, h2, w)[0]))
    for got, ref in losses:
        assert abs(got - ref) <= TOL * abs(ref), losses
    assert losses[0][0] == losses[2][0]


def test_argument_errors():
    from ssl_audio_b200.projector import proj_tail_fwd
    pack = torch.zeros(7 * 64, device="cuda")
    h = torch.zeros(8, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        proj_tail_fwd(h.float(), h, torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16), pack)
    with pytest.raises(ValueError):
        proj_tail_fwd(h, h, torch.zeros(64, 32, device="cuda", dtype=torch.bfloat16), pack)
    with pytest.raises(ValueError):
        proj_tail_fwd(h[:, :60].contiguous(), h[:, :60].contiguous(), torch.zeros(64, 60, device="cuda", dtype=torch.bfloat16), pack)     # K % 8
    with pytest.raises(RuntimeError):
        proj_tail_fwd(h.cpu(), h.cpu(), torch.zeros(64, 64, dtype=torch.bfloat16), pack)
