"""World-size-2 test of the multi-GPU objective's host logic (ssl_audio_b200/dist.py) on CPUs with gloo: dimension
partition, all-gather order, gradient all-to-all and loss reduction.  The row-block compute is injected from the
numpy oracle here (test infrastructure); on a GPU box the same choreography drives abt_bt_loss_rows_fwd_bwd
(tests/test_gpu_dist.py)."""
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


def _oracle_rows_fn(zg1, zg2, begin, count, alpha, lmbda, hsic, eps, momentum, grad_scale, need_mask, running_mean, running_var):
    from oracle import abt_oracle as O
    z1, z2 = zg1.float().numpy(), zg2.float().numpy()
    loss, dz1, dz2, c = O.bt_loss_forward_backward(z1, z2, alpha, lmbda, hsic, eps)
    blk = c[begin:begin + count].copy()
    idx = np.arange(begin, begin + count)
    blk[idx - begin, idx] = 0.0
    parts = torch.tensor([(blk ** 2).sum(), blk.sum(), ((np.diagonal(c) - 1.0) ** 2).sum()], dtype=torch.float64)
    d1 = torch.from_numpy(dz1[:, begin:begin + count] * grad_scale).to(zg1.dtype) if need_mask & 1 else None
    d2 = torch.from_numpy(dz2[:, begin:begin + count] * grad_scale).to(zg1.dtype) if need_mask & 2 else None
    return parts, d1, d2


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
        loss, dz1, dz2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, rows_fn=_oracle_rows_fn)
        ref_loss, r1, r2, _ = O.bt_loss_forward_backward(z1g, z2g, 1.0, 0.005, hsic)
        ok = abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
        # default grad_scale = world_size (cancels DDP's averaging)
        ok = ok and np.allclose(dz1.numpy(), world * r1[rank * n:(rank + 1) * n], rtol=1e-4, atol=1e-7)
        ok = ok and np.allclose(dz2.numpy(), world * r2[rank * n:(rank + 1) * n], rtol=1e-4, atol=1e-7)
        loss2, e1, e2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, rows_fn=_oracle_rows_fn, need_dz1=False, grad_scale=1.0)
        ok = ok and e1 is None and np.allclose(e2.numpy(), r2[rank * n:(rank + 1) * n], rtol=1e-4, atol=1e-7)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("hsic", [False, True])
def test_global_objective_world2_gloo(hsic):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), hsic, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_row_block_partition():
    from ssl_audio_b200.dist import row_block
    assert [row_block(8192, 8, r) for r in range(8)] == [(r * 1024, 1024) for r in range(8)]
    assert row_block(256, 8, 3) == (96, 32)
    assert row_block(64, 1, 0) == (0, 64)
    blocks = [row_block(2048, 4, r) for r in range(4)]
    assert sum(c for _, c in blocks) == 2048 and all(b % 8 == 0 and c % 8 == 0 for b, c in blocks)
