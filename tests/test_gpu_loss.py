"""GPU parity of the Barlow Twins objective (through the C ABI) against the numpy oracle and the
reference-generated golden fixtures.  Tolerances are those of BASELINE.json:north_star: loss and
gradients within 1e-3 relative (gradients: Frobenius-norm relative error) for bf16 inputs with
fp32 accumulation; the oracle consumes the same bf16-rounded values in float64."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402

TOL = 1e-3


def _cfg(d, hsic=False, alpha=1.0, lmbda=0.005):
    return types.SimpleNamespace(projector_out_dim=d, HSIC=hsic, alpha=alpha, lmbda=lmbda)


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def _run(z1, z2, dtype, hsic=False, alpha=1.0, lmbda=0.005, need=(True, True)):
    import ssl_audio_b200 as S
    t1 = torch.from_numpy(z1).cuda().to(dtype).requires_grad_(need[0])
    t2 = torch.from_numpy(z2).cuda().to(dtype).requires_grad_(need[1])
    mod = S.BarlowTwinsLoss(_cfg(z1.shape[1], hsic, alpha, lmbda), ncrops=2).cuda()
    loss = mod.forward_loss(t1, t2)
    loss.backward()
    torch.cuda.synchronize()
    g1 = t1.grad.float().cpu().numpy() if need[0] else None
    g2 = t2.grad.float().cpu().numpy() if need[1] else None
    return float(loss), g1, g2, mod


@pytest.mark.parametrize("n,d", [(32, 64), (128, 256), (48, 320), (128, 2048), (256, 1024), (100, 512)])
def test_bf16_matches_oracle(n, d):
    z1, z2 = O.synth_embeddings(n, d, seed=n + d)
    loss, g1, g2, _ = _run(z1, z2, torch.bfloat16)
    rl, r1, r2, _ = O.bt_loss_forward_backward(z1, z2)
    assert abs(loss - rl) <= TOL * abs(rl), (loss, rl)
    # outputs are rounded to bf16 (gradients in the input dtype): allow the bf16 quantum on top
    assert _rel(g1, r1) < TOL + 4e-3, _rel(g1, r1)
    assert _rel(g2, r2) < TOL + 4e-3, _rel(g2, r2)


@pytest.mark.parametrize("n,d,hsic", [(64, 256, False), (64, 256, True), (128, 1024, False)])
def test_fp32_io_matches_oracle_on_bf16_rounded_inputs(n, d, hsic):
    z1, z2 = O.synth_embeddings(n, d, seed=7)          # already bf16-representable
    loss, g1, g2, _ = _run(z1, z2, torch.float32, hsic=hsic)
    rl, r1, r2, _ = O.bt_loss_forward_backward(z1, z2, hsic=hsic)
    assert abs(loss - rl) <= TOL * abs(rl), (loss, rl)
    assert _rel(g1, r1) < TOL, _rel(g1, r1)
    assert _rel(g2, r2) < TOL, _rel(g2, r2)


def test_large_n_and_shifted_means():
    # N > 256 exercises the n-tiling of the gradient GEMM; a large common offset exercises the
    # rank-1 batch-norm correction (|mu| >> sigma)
    z1, z2 = O.synth_embeddings(640, 512, seed=3)
    z1 = O.round_bf16(z1 + 3.0)
    z2 = O.round_bf16(z2 * 0.25 - 2.0)
    loss, g1, g2, _ = _run(z1, z2, torch.float32)
    rl, r1, r2, _ = O.bt_loss_forward_backward(z1, z2)
    assert abs(loss - rl) <= TOL * abs(rl), (loss, rl)
    assert _rel(g1, r1) < TOL and _rel(g2, r2) < TOL, (_rel(g1, r1), _rel(g2, r2))


def test_stop_gradient_side_is_skipped():
    z1, z2 = O.synth_embeddings(64, 256, seed=5)
    loss, g1, g2, _ = _run(z1, z2, torch.float32, need=(False, True))
    rl, _, r2, _ = O.bt_loss_forward_backward(z1, z2)
    assert g1 is None
    assert abs(loss - rl) <= TOL * abs(rl) and _rel(g2, r2) < TOL


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_reference_outputs(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    z1, z2, hsic = g[f"{tag}_z1"], g[f"{tag}_z2"], bool(g[f"{tag}_hsic"])
    loss, g1, g2, mod = _run(z1, z2, torch.float32, hsic=hsic)
    ref = float(g[f"{tag}_loss"])
    assert abs(loss - ref) <= TOL * abs(ref), (loss, ref)
    assert _rel(g1, g[f"{tag}_dz1"].astype(np.float64)) < TOL
    assert _rel(g2, g[f"{tag}_dz2"].astype(np.float64)) < TOL
    sd = mod.state_dict()
    assert set(sd.keys()) == {"bn.running_mean", "bn.running_var", "bn.num_batches_tracked"}
    np.testing.assert_allclose(sd["bn.running_mean"].cpu().numpy(), g[f"{tag}_running_mean"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sd["bn.running_var"].cpu().numpy(), g[f"{tag}_running_var"], rtol=1e-4, atol=1e-5)
    assert int(sd["bn.num_batches_tracked"]) == int(g[f"{tag}_num_batches"])


def test_byol_pairing_two_terms(golden_dir):
    import ssl_audio_b200 as S
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    s = torch.from_numpy(g["byol_student"]).cuda().requires_grad_(True)
    t = torch.from_numpy(g["byol_teacher"]).cuda().requires_grad_(True)
    mod = S.BarlowTwinsLoss(_cfg(64), ncrops=2).cuda()
    loss = mod(s, t, ngcrops_each=2)
    loss.backward()
    ref = float(g["byol_loss"])
    assert abs(float(loss) - ref) <= TOL * abs(ref)
    assert _rel(s.grad.cpu().numpy(), g["byol_dstudent"].astype(np.float64)) < TOL
    assert _rel(t.grad.cpu().numpy(), g["byol_dteacher"].astype(np.float64)) < TOL


def test_grad_output_scaling_and_fp16():
    z1, z2 = O.synth_embeddings(64, 256, seed=9)
    import ssl_audio_b200 as S
    t1 = torch.from_numpy(z1).cuda().half().requires_grad_(True)
    t2 = torch.from_numpy(z2).cuda().half().requires_grad_(True)
    mod = S.BarlowTwinsLoss(_cfg(256), ncrops=2).cuda()
    (mod.forward_loss(t1, t2) * 128.0).backward()     # GradScaler-style scaling
    _, r1, r2, _ = O.bt_loss_forward_backward(z1, z2)
    assert _rel(t1.grad.float().cpu().numpy() / 128.0, r1) < 5e-3
    assert _rel(t2.grad.float().cpu().numpy() / 128.0, r2) < 5e-3


def test_error_conventions():
    import ssl_audio_b200 as S
    mod = S.BarlowTwinsLoss(_cfg(64), ncrops=2).cuda()
    with pytest.raises(RuntimeError):
        mod.forward_loss(torch.zeros(8, 64), torch.zeros(8, 64))            # CPU tensors: no fallback
    with pytest.raises(ValueError):
        mod.forward_loss(torch.zeros(8, 64).cuda(), torch.zeros(4, 64).cuda())
    with pytest.raises(ValueError):
        S.BarlowTwinsLoss(_cfg(72), ncrops=2).cuda().forward_loss(torch.zeros(8, 72).cuda(), torch.zeros(8, 72).cuda())
    with pytest.raises(AssertionError):
        S.off_diagonal(torch.zeros(3, 4))


def test_full_size_properties_d8192():
    """BASELINE config 3 at full size (N=128, D=8192): properties that need no O(D^2) oracle --
    loss(z, z) has a vanishing on-diagonal term, the loss is symmetric in its arguments, gradients
    are orthogonal to the batch-norm null space (column sums and column projections on zh vanish)."""
    n, d = 128, 8192
    z1, z2 = O.synth_embeddings(n, d, seed=1)
    loss12, g1, g2, _ = _run(z1, z2, torch.bfloat16)
    loss21, h2, h1, _ = _run(z2, z1, torch.bfloat16)
    assert abs(loss12 - loss21) <= 1e-5 * abs(loss12)
    assert _rel(g1, h1.astype(np.float64)) < 1e-2 and _rel(g2, h2.astype(np.float64)) < 1e-2
    col_sum = np.abs(g1.astype(np.float64).sum(0)).max()
    assert col_sum < 1e-2 * np.abs(g1).sum(0).max()
    # spot-check 64 columns of dz1 against the closed form evaluated only for those columns
    rl, r1, _, _ = O.bt_loss_forward_backward(z1[:, :], z2[:, :]) if False else (None, None, None, None)
    h1z, _, _, rr1 = O.batchnorm_train(z1.astype(np.float64))
    h2z, _, _, _ = O.batchnorm_train(z2.astype(np.float64))
    cols = np.arange(0, d, d // 64)
    c_rows = h1z[:, cols].T @ h2z / n                                   # (64, D) rows of C
    G = 2 * 0.005 * c_rows
    G[np.arange(len(cols)), cols] = 2 * (c_rows[np.arange(len(cols)), cols] - 1.0)
    gh = h2z @ G.T / n                                                   # (N, 64)
    hz = h1z[:, cols]
    ref = (gh - gh.mean(0) - hz * (gh * hz).mean(0)) * rr1[cols]
    assert _rel(g1[:, cols], ref) < TOL + 4e-3
    # loss against a float64 evaluation that never forms more than a row block of C
    on = 0.0
    off = 0.0
    for s in range(0, d, 1024):
        blk = h1z[:, s:s + 1024].T @ h2z / n
        idx = np.arange(s, min(s + 1024, d))
        dg = blk[idx - s, idx]
        on += ((dg - 1) ** 2).sum()
        off += (blk ** 2).sum() - (dg ** 2).sum()
    ref_loss = on + 0.005 * off
    assert abs(loss12 - ref_loss) <= TOL * ref_loss, (loss12, ref_loss)


def test_bench_size_properties_n1024_d8192():
    """The bench workload (N = 1024 rows, D = 8192, bf16): same size-independent checks as above -- symmetry of the loss under swapping the
    views, gradients in the batch-norm null space, 64 spot-checked gradient columns of BOTH views against the closed form, and the loss
    against a float64 evaluation done row block by row block."""
    n, d = 1024, 8192
    z1, z2 = O.synth_embeddings(n, d, seed=2)
    loss12, g1, g2, _ = _run(z1, z2, torch.bfloat16)
    loss21, h2, h1, _ = _run(z2, z1, torch.bfloat16)
    assert abs(loss12 - loss21) <= 1e-5 * abs(loss12)
    assert _rel(g1, h1.astype(np.float64)) < 1e-2 and _rel(g2, h2.astype(np.float64)) < 1e-2
    assert np.abs(g1.astype(np.float64).sum(0)).max() < 1e-2 * np.abs(g1).sum(0).max()
    h1z, _, _, rr1 = O.batchnorm_train(z1.astype(np.float64))
    h2z, _, _, rr2 = O.batchnorm_train(z2.astype(np.float64))
    cols = np.arange(3, d, d // 64)
    ar = np.arange(len(cols))
    # dz1[:, cols]: rows `cols` of C
    c_rows = h1z[:, cols].T @ h2z / n
    G = 2 * 0.005 * c_rows
    G[ar, cols] = 2 * (c_rows[ar, cols] - 1.0)
    gh = h2z @ G.T / n
    ref1 = (gh - gh.mean(0) - h1z[:, cols] * (gh * h1z[:, cols]).mean(0)) * rr1[cols]
    assert _rel(g1[:, cols], ref1) < TOL + 4e-3
    # dz2[:, cols]: columns `cols` of C
    c_cols = h1z.T @ h2z[:, cols] / n
    G = 2 * 0.005 * c_cols
    G[cols, ar] = 2 * (c_cols[cols, ar] - 1.0)
    gh = h1z @ G / n
    ref2 = (gh - gh.mean(0) - h2z[:, cols] * (gh * h2z[:, cols]).mean(0)) * rr2[cols]
    assert _rel(g2[:, cols], ref2) < TOL + 4e-3
    on = off = 0.0
    for s in range(0, d, 1024):
        blk = h1z[:, s:s + 1024].T @ h2z / n
        idx = np.arange(s, min(s + 1024, d))
        dg = blk[idx - s, idx]
        on += ((dg - 1) ** 2).sum()
        off += (blk ** 2).sum() - (dg ** 2).sum()
    ref_loss = on + 0.005 * off
    assert abs(loss12 - ref_loss) <= TOL * ref_loss, (loss12, ref_loss)
