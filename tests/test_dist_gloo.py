"""World-size-2 and -4 tests of the multi-GPU objective's host logic (ssl_audio_b200/dist.py) on CPUs with gloo: dimension
partition, all-gather order, gradient all-to-all and loss reduction.  The three compute stages are injected as a
numpy backend here (test infrastructure); on a GPU box the same choreography drives abt_bt_dist_stats_local /
abt_bt_dist_normalize / abt_bt_dist_rows_fwd_bwd (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _NumpyBackend:
    """CPU stand-in for ssl_audio_b200.dist.CudaBackend (test infrastructure): the same three stages -- local statistics pack,
    global statistics + fp16 standardised local rows, row block of the objective on the gathered standardised rows -- in numpy
    float64, so that the collective choreography of dist.py can run under gloo."""

    def workspace(self, device, n_local, world, d, row_count):
        ng = n_local * world
        return dict(zh1=torch.zeros((ng, d), dtype=torch.float16), zh2=torch.zeros((ng, d), dtype=torch.float16),
                    pack_local=torch.zeros(7 * d), pack_all=torch.zeros(world * 7 * d), stats={})

    def stats_local(self, w, z1, z2, world, row_count):
        a, b = z1.double().numpy(), z2.double().numpy()
        k1, k2 = a[0], b[0]
        da, db = a - k1, b - k2
        pack = np.stack([da.sum(0), (da * da).sum(0), db.sum(0), (db * db).sum(0), (da * db).sum(0), k1, k2])
        w["pack_local"].copy_(torch.from_numpy(pack.reshape(-1)).float())

    def normalize(self, w, z1, z2, world, rank, row_count, eps, momentum, running_mean, running_var):
        n, d = z1.shape
        pk = w["pack_all"].double().numpy().reshape(world, 7, d)
        ng = n * world
        m1 = (pk[:, 0] + n * pk[:, 5]).sum(0) / ng
        m2 = (pk[:, 2] + n * pk[:, 6]).sum(0) / ng
        d1, d2 = pk[:, 5] - m1, pk[:, 6] - m2
        v1 = (pk[:, 1] + 2 * d1 * pk[:, 0] + n * d1 * d1).sum(0) / ng
        v2 = (pk[:, 3] + 2 * d2 * pk[:, 2] + n * d2 * d2).sum(0) / ng
        cv = (pk[:, 4] + d2 * pk[:, 0] + d1 * pk[:, 2] + n * d1 * d2).sum(0) / ng
        r1, r2 = 1 / np.sqrt(v1 + eps), 1 / np.sqrt(v2 + eps)
        w["stats"] = dict(r1=r1, r2=r2, cdiag=cv * r1 * r2)
        w["zh1"][rank * n:(rank + 1) * n] = torch.from_numpy((z1.double().numpy() - m1) * r1).half()
        w["zh2"][rank * n:(rank + 1) * n] = torch.from_numpy((z2.double().numpy() - m2) * r2).half()

    def rows(self, w, dtype, device, n_local, world, d, begin, count, alpha, lmbda, hsic, grad_scale, need_mask, phase=0):
        h1, h2 = w["zh1"].double().numpy(), w["zh2"].double().numpy()
        ng = h1.shape[0]
        st = w["stats"]
        c = h1.T @ h2 / ng
        np.fill_diagonal(c, st["cdiag"])                       # the diagonal comes from the fp32/fp64 statistics, as on the GPU
        g = 2 * lmbda * (c + (1.0 if hsic else 0.0))
        np.fill_diagonal(g, 2 * alpha * (np.diagonal(c) - 1.0))
        idx = np.arange(begin, begin + count)
        blk = c[idx].copy()
        blk[idx - begin, idx] = 0.0
        parts = torch.tensor([(blk ** 2).sum(), blk.sum(), ((st["cdiag"] - 1.0) ** 2).sum()], dtype=torch.float64)

        def bn_bwd(gh, hz, r):
            return (gh - gh.mean(0) - hz * (gh * hz).mean(0)) * r
        d1 = d2 = None
        if (need_mask & 1) and phase != 2:
            gh = h2 @ g[idx].T / ng
            d1 = torch.from_numpy(bn_bwd(gh, h1[:, idx], st["r1"][idx]) * grad_scale).to(dtype)
        if (need_mask & 2) and phase != 1:
            gh = h1 @ g[:, idx] / ng
            d2 = torch.from_numpy(bn_bwd(gh, h2[:, idx], st["r2"][idx]) * grad_scale).to(dtype)
        return (parts if phase != 2 else None), d1, d2


def _worker(rank, world, port, hsic, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import abt_oracle as O
        from ssl_audio_b200 import dist as D
        n, d = 12, 64
        z1g, z2g = O.synth_embeddings(world * n, d, seed=5)
        z1 = torch.from_numpy(z1g[rank * n:(rank + 1) * n])
        z2 = torch.from_numpy(z2g[rank * n:(rank + 1) * n])
        assert D.is_active()
        be = _NumpyBackend()
        loss, dz1, dz2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, backend=be)
        ref_loss, r1, r2, _ = O.bt_loss_forward_backward(z1g, z2g, 1.0, 0.005, hsic)
        ok = abs(float(loss) - ref_loss) <= 1e-3 * abs(ref_loss)          # standardised rows travel in fp16

        def rel(a, b):
            return np.linalg.norm(a - b) / np.linalg.norm(b)
        # default grad_scale = world_size (cancels DDP's averaging)
        ok = ok and rel(dz1.numpy(), world * r1[rank * n:(rank + 1) * n]) < 2e-3
        ok = ok and rel(dz2.numpy(), world * r2[rank * n:(rank + 1) * n]) < 2e-3
        loss2, e1, e2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, backend=be, need_dz1=False, grad_scale=1.0)
        ok = ok and e1 is None and rel(e2.numpy(), r2[rank * n:(rank + 1) * n]) < 2e-3
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,hsic", [(2, False), (2, True), (4, False)])
def test_global_objective_gloo(world, hsic):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), hsic, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def test_row_block_partition():
    from ssl_audio_b200.dist import row_block
    assert [row_block(8192, 8, r) for r in range(8)] == [(r * 1024, 1024) for r in range(8)]
    assert row_block(256, 8, 3) == (96, 32)
    assert row_block(64, 1, 0) == (0, 64)
    blocks = [row_block(2048, 4, r) for r in range(4)]
    assert sum(c for _, c in blocks) == 2048 and all(b % 8 == 0 and c % 8 == 0 for b, c in blocks)
