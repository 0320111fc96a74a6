"""Row-block (multi-GPU) form of the objective on ONE GPU: the R ranks' calls are issued one after the other on
the same gathered batch and their outputs assembled, which is exactly what the collectives in
ssl_audio_b200/dist.py do across GPUs (that choreography is tested with gloo in tests/test_dist_gloo.py).
Parity target: the single-process global-batch oracle ("Multi-GPU")."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import abt_oracle as O  # noqa: E402

TOL = 1e-3
ROUND_TOL = 2e-4      # 16-bit outputs against the rounding of the fp32 outputs of the same kernels


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def _assemble_row_blocks(world, n_local, d, hsic, tdt, z1, z2):
    from ssl_audio_b200 import dist as D
    ng = world * n_local
    zg1, zg2 = torch.from_numpy(z1).cuda().to(tdt), torch.from_numpy(z2).cuda().to(tdt)
    off = np.zeros(2)
    on = None
    dz1 = np.zeros((ng, d), np.float32)
    dz2 = np.zeros((ng, d), np.float32)
    rm = [torch.zeros(d, device="cuda") for _ in range(world)]
    rv = [torch.ones(d, device="cuda") for _ in range(world)]
    for r in range(world):
        begin, count = D.row_block(d, world, r)
        parts, a, b = D._rows_cuda(zg1, zg2, begin, count, 1.0, 0.005, hsic, 1e-5, 0.1, 1.0, 3, rm[r], rv[r])
        p = parts.cpu().numpy()
        off += p[:2]
        on = p[2] if on is None else on
        assert abs(p[2] - on) <= 1e-9 * max(1.0, abs(on))          # identical on every rank
        dz1[:, begin:begin + count] = a.float().cpu().numpy()
        dz2[:, begin:begin + count] = b.float().cpu().numpy()
    loss = on + 0.005 * (off[0] + (2 * off[1] + d * (d - 1) if hsic else 0.0))
    return loss, dz1, dz2, rm, rv


def _check_against_oracle(run, dtype, z1, z2, hsic, world, d):
    """fp32 outputs within 1e-3 of the oracle; bf16 outputs = the rounding of the fp32 ones (no widened tolerance)."""
    rl, r1, r2, _ = O.bt_loss_forward_backward(z1, z2, 1.0, 0.005, hsic)
    loss, dz1, dz2, rm, rv = run(torch.float32)
    assert abs(loss - rl) <= TOL * abs(rl), (loss, rl)
    assert _rel(dz1, r1) < TOL and _rel(dz2, r2) < TOL, (_rel(dz1, r1), _rel(dz2, r2))
    if dtype == "bf16":
        loss_b, b1, b2, rm, rv = run(torch.bfloat16)
        assert abs(loss_b - rl) <= TOL * abs(rl), (loss_b, rl)
        for b, a in ((b1, dz1), (b2, dz2)):
            rounded = torch.from_numpy(a).bfloat16().float().numpy().astype(np.float64)
            assert _rel(b, rounded) < ROUND_TOL, _rel(b, rounded)
    # running statistics are those of the global batch on every rank
    m, v = O.bn_running_update(np.zeros(d), np.ones(d), z1)
    m, v = O.bn_running_update(m, v, z2)
    for r in range(world):
        np.testing.assert_allclose(rm[r].cpu().numpy(), m, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(rv[r].cpu().numpy(), v, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("world,n_local,d,hsic,dtype", [
    (2, 64, 256, False, "f32"), (4, 32, 512, False, "f32"), (8, 16, 1024, True, "f32"), (2, 160, 512, False, "bf16"),
    (8, 128, 2048, False, "bf16"),
])
def test_row_blocks_assemble_to_global_objective(world, n_local, d, hsic, dtype):
    z1, z2 = O.synth_embeddings(world * n_local, d, seed=world + d)
    _check_against_oracle(lambda tdt: _assemble_row_blocks(world, n_local, d, hsic, tdt, z1, z2), dtype, z1, z2, hsic, world, d)


def test_row_block_one_sided_and_errors():
    from ssl_audio_b200 import dist as D
    z1, z2 = O.synth_embeddings(64, 256, seed=2)
    zg1, zg2 = torch.from_numpy(z1).cuda(), torch.from_numpy(z2).cuda()
    parts, a, b = D._rows_cuda(zg1, zg2, 128, 64, 1.0, 0.005, False, 1e-5, 0.1, 2.0, 2, None, None)
    assert a is None
    _, _, r2, _ = O.bt_loss_forward_backward(z1, z2)
    assert _rel(b.cpu().numpy(), 2.0 * r2[:, 128:192]) < TOL        # grad_scale = 2
    with pytest.raises(ValueError):
        D._rows_cuda(zg1, zg2, 4, 64, 1.0, 0.005, False, 1e-5, 0.1, 1.0, 3, None, None)     # unaligned block
    with pytest.raises(ValueError):
        D._rows_cuda(zg1, zg2, 224, 64, 1.0, 0.005, False, 1e-5, 0.1, 1.0, 3, None, None)   # outside D


@pytest.mark.parametrize("world,n_local,d,hsic,dtype", [
    (2, 64, 256, False, "f32"), (4, 32, 512, True, "f32"), (8, 128, 2048, False, "bf16"), (2, 100, 320, False, "f32"),
])
def test_three_stage_pipeline_emulated_on_one_gpu(world, n_local, d, hsic, dtype):
    """abt_bt_dist_stats_local -> [all-gather] -> abt_bt_dist_normalize -> [all-gather] -> abt_bt_dist_rows_fwd_bwd with the R ranks
    emulated one after the other (one workspace per rank, the collectives done by hand)."""
    if d % (8 * world) != 0:
        pytest.skip("D must split into 8-aligned blocks")
    ng = world * n_local
    z1, z2 = O.synth_embeddings(ng, d, seed=3 * world + d)
    z1 = O.round_bf16(z1 + 0.5)                      # shifted means: exercises the re-centring of the combined sums
    _check_against_oracle(lambda tdt: _three_stage(world, n_local, d, hsic, tdt, z1, z2), dtype, z1, z2, hsic, world, d)


def _three_stage(world, n_local, d, hsic, tdt, z1, z2):
    from ssl_audio_b200 import dist as D
    ng = world * n_local
    zg1, zg2 = torch.from_numpy(z1).cuda().to(tdt), torch.from_numpy(z2).cuda().to(tdt)
    count = D.row_block(d, world, 0)[1]
    bes = [D.CudaBackend() for _ in range(world)]
    wss = [bes[r].workspace(zg1.device, n_local, world, d, count) for r in range(world)]
    loc = [(zg1[r * n_local:(r + 1) * n_local].contiguous(), zg2[r * n_local:(r + 1) * n_local].contiguous()) for r in range(world)]
    for r in range(world):
        bes[r].stats_local(wss[r], loc[r][0], loc[r][1], world, count)
    packs = torch.cat([wss[r]["pack_local"] for r in range(world)])
    rm = [torch.zeros(d, device="cuda") for _ in range(world)]
    rv = [torch.ones(d, device="cuda") for _ in range(world)]
    for r in range(world):
        wss[r]["pack_all"].copy_(packs)
        bes[r].normalize(wss[r], loc[r][0], loc[r][1], world, r, count, 1e-5, 0.1, rm[r], rv[r])
    for r in range(world):                            # the all-gather of the standardised rows
        for q in range(world):
            if q != r:
                wss[r]["zh1"][q * n_local:(q + 1) * n_local].copy_(wss[q]["zh1"][q * n_local:(q + 1) * n_local])
                wss[r]["zh2"][q * n_local:(q + 1) * n_local].copy_(wss[q]["zh2"][q * n_local:(q + 1) * n_local])
    off = np.zeros(2)
    on = None
    dz1 = np.zeros((ng, d), np.float32)
    dz2 = np.zeros((ng, d), np.float32)
    for r in range(world):
        begin, cnt = D.row_block(d, world, r)
        if r % 2 == 0:                                  # one launch ...
            parts, a, b = bes[r].rows(wss[r], tdt, zg1.device, n_local, world, d, begin, cnt, 1.0, 0.005, hsic, 1.0, 3)
        else:                                           # ... or the two-phase form the host uses to overlap the gradient all-to-all
            parts, a, _ = bes[r].rows(wss[r], tdt, zg1.device, n_local, world, d, begin, cnt, 1.0, 0.005, hsic, 1.0, 3, 1)
            _, _, b = bes[r].rows(wss[r], tdt, zg1.device, n_local, world, d, begin, cnt, 1.0, 0.005, hsic, 1.0, 3, 2)
        p = parts.cpu().numpy()
        off += p[:2]
        on = p[2] if on is None else on
        assert abs(p[2] - on) <= 1e-9 * max(1.0, abs(on))
        dz1[:, begin:begin + cnt] = a.float().cpu().numpy()
        dz2[:, begin:begin + cnt] = b.float().cpu().numpy()
    loss = on + 0.005 * (off[0] + (2 * off[1] + d * (d - 1) if hsic else 0.0))
    return loss, dz1, dz2, rm, rv


def test_native_multi_gpu_step_under_torchrun():
    """Two real ranks over NCCL (skipped on a single-GPU box): the native one-call step and the torch.distributed choreography
    against the global-batch oracle (tools/dist_check.py)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_side_stream_hook_is_gated_on_the_objective_stream():
    """The comm-overlap hook fires on the host long before the device reaches the hook point: work it puts on a side stream must
    wait for the objective's stream (ssl_audio_b200.dist.side_stream_hook), and must run on the side stream."""
    from ssl_audio_b200.dist import side_stream_hook
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(dev)
    x = torch.zeros(1, device=dev)
    seen = {}

    def fn():
        seen["stream"] = torch.cuda.current_stream(dev)
        seen["y"] = x.clone()                      # runs on the side stream

    hook = side_stream_hook(side, fn)
    torch.cuda.synchronize()
    torch.cuda._sleep(200_000_000)                 # ~0.1 s of device time on the objective's stream ...
    x.fill_(1.0)                                   # ... then the value the side stream must see
    hook()
    torch.cuda.synchronize()
    assert seen["stream"] == side
    assert float(seen["y"]) == 1.0
