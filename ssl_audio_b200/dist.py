"""Multi-GPU Barlow Twins objective on the GLOBAL batch (one process per GPU, torch.distributed / NCCL).

Reference behaviour being replaced: every rank builds its own D x D matrix from its LOCAL batch (local,
unsynchronised BatchNorm) and the ranks `all_reduce(SUM)` those matrices (utils/loss.py:17-21) -- 256 MiB per
rank at D = 8192, and for R > 1 the result is R times the mean of per-rank correlation matrices (SURVEY.md
section 3.3, quirk 1).  North-star semantics implemented here: ONE objective over the global batch of
N_g = sum of local batches, identical (up to rounding) to the single-process reference run on the
rank-ordered concatenation of the batches:

    1. all-gather the raw bf16 embeddings of both views                       (NCCL, 2 x N_g x D x 2 B)
    2. every rank derives the global column statistics from the gathered data   (redundant, bandwidth-trivial)
    3. rank r computes rows [r D/R, (r+1) D/R) of C and of C^T on the tensor cores, its share of the
       off-diagonal loss, and d loss / d z for ALL samples restricted to its dimensions
       (abt_bt_loss_rows_fwd_bwd; batch-norm backward is per column, so no partial sums cross ranks)
    4. all-to-all returns each sample's gradient slice to the rank that owns the sample (NCCL)
    5. a 2-double all-reduce completes the loss

`grad_scale`: DDP averages parameter gradients over ranks, so the local d loss / d z is multiplied by
`world_size` by default; single-process and R-rank runs then produce the same parameter update.

The collective choreography is separated from the row-block compute (`rows_fn`) so that it can be exercised
with the gloo backend on CPUs (tests/test_dist_gloo.py passes an oracle-backed `rows_fn`); the product default
is the CUDA entry point and nothing else.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

_DTYPES = {torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


_WS = {}


def _rows_cuda(zg1: torch.Tensor, zg2: torch.Tensor, row_begin: int, row_count: int, alpha: float, lmbda: float, hsic: bool,
               eps: float, momentum: float, grad_scale: float, need_mask: int, running_mean, running_var):
    """Row-block objective on the GPU.  Returns (parts (3,) float64 device, dzr1, dzr2) with dzr* (N_g, row_count)."""
    if not zg1.is_cuda:
        raise RuntimeError("embeddings must be CUDA tensors: ssl_audio_b200 has no CPU path")
    lib = _lib.load()
    n, d = int(zg1.shape[0]), int(zg1.shape[1])
    code = _DTYPES[zg1.dtype]
    dev = zg1.device
    key = (dev.index, n, d, row_count, code)
    if key not in _WS:
        nbytes = C.c_size_t()
        _lib.check(lib.abt_bt_rows_workspace_bytes(n, d, row_count, code, C.byref(nbytes)))
        _WS[key] = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
    buf = _WS[key]
    parts = torch.empty(3, dtype=torch.float64, device=dev)
    dzr1 = torch.empty((n, row_count), dtype=zg1.dtype, device=dev) if need_mask & 1 else None
    dzr2 = torch.empty((n, row_count), dtype=zg1.dtype, device=dev) if need_mask & 2 else None
    a = _lib.BtRowsArgs()
    a.zg1, a.zg2 = zg1.data_ptr(), zg2.data_ptr()
    a.dtype, a.n_rows, a.n_dims = code, n, d
    a.row_begin, a.row_count = int(row_begin), int(row_count)
    a.alpha, a.lambda_, a.hsic = float(alpha), float(lmbda), int(bool(hsic))
    a.eps, a.momentum, a.grad_scale = float(eps), float(momentum), float(grad_scale)
    a.need_grad_mask = int(need_mask)
    a.loss_parts = parts.data_ptr()
    a.dzr1 = dzr1.data_ptr() if dzr1 is not None else None
    a.dzr2 = dzr2.data_ptr() if dzr2 is not None else None
    a.running_mean = running_mean.data_ptr() if running_mean is not None else None
    a.running_var = running_var.data_ptr() if running_var is not None else None
    a.workspace = (buf.data_ptr() + 255) // 256 * 256
    a.workspace_bytes = buf.numel() - 256
    with torch.cuda.device(dev):
        _lib.check(lib.abt_bt_loss_rows_fwd_bwd(C.byref(a), torch.cuda.current_stream(dev).cuda_stream))
    return parts, dzr1, dzr2


def row_block(d: int, world: int, rank: int) -> Tuple[int, int]:
    """Dimensions owned by `rank`: contiguous blocks of ceil(D / R) rounded up to 8 (the last block may be shorter)."""
    per = -(-d // world)
    per = -(-per // 8) * 8
    begin = min(rank * per, d)
    return begin, max(0, min(per, d - begin))


def bt_loss_fwd_bwd_global(z1: torch.Tensor, z2: torch.Tensor, alpha: float, lmbda: float, hsic: bool, *, eps: float = 1e-5,
                           momentum: float = 0.1, running_mean: Optional[torch.Tensor] = None,
                           running_var: Optional[torch.Tensor] = None, need_dz1: bool = True, need_dz2: bool = True,
                           grad_scale: Optional[float] = None, group=None, rows_fn: Optional[Callable] = None):
    """Global-batch Barlow Twins loss + local gradients.  z1, z2: this rank's (N, D) embeddings (same N on every rank).
    Returns (loss 0-dim fp32, identical on every rank; dz1 (N, D) or None; dz2 (N, D) or None)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows_fn = rows_fn or _rows_cuda
    if grad_scale is None:
        grad_scale = float(world)
    z1 = z1.contiguous()
    z2 = z2.contiguous()
    n, d = int(z1.shape[0]), int(z1.shape[1])
    if d % 8 != 0:
        raise ValueError("projector_out_dim must be a multiple of 8")
    begin, count = row_block(d, world, rank)
    per = row_block(d, world, 0)[1]
    if count != per:
        raise ValueError(f"D = {d} does not split into equal 8-aligned blocks over {world} ranks")
    # 1. all-gather the raw embeddings (rank-ordered concatenation = the single-process global batch)
    zg1 = torch.empty((world * n, d), dtype=z1.dtype, device=z1.device)
    zg2 = torch.empty((world * n, d), dtype=z2.dtype, device=z2.device)
    dist.all_gather_into_tensor(zg1, z1, group=group)
    dist.all_gather_into_tensor(zg2, z2, group=group)
    # 2.+3. statistics, row block of C / C^T, gradients of all samples for the dimensions of this rank
    need_mask = (1 if need_dz1 else 0) | (2 if need_dz2 else 0)
    parts, dzr1, dzr2 = rows_fn(zg1, zg2, begin, count, alpha, lmbda, hsic, eps, momentum, grad_scale, need_mask, running_mean,
                                running_var)
    # 4. all-to-all: (R, N, D/R) blocks of my dimensions go to the ranks owning the samples
    def exchange(dzr):
        if dzr is None:
            return None
        recv = torch.empty((world, n, count), dtype=dzr.dtype, device=dzr.device)
        dist.all_to_all_single(recv, dzr.view(world, n, count), group=group)
        return recv.permute(1, 0, 2).reshape(n, d)        # [src rank = dimension block][n] -> (n, D)
    dz1 = exchange(dzr1)
    dz2 = exchange(dzr2)
    # 5. loss: the off-diagonal partial sums are per row block, the on-diagonal sum is already global
    off = parts[:2].clone()
    dist.all_reduce(off, group=group)
    on = parts[2]
    off_total = off[0] + (2.0 * off[1] + float(d) * float(d - 1) if hsic else 0.0)
    loss = (alpha * on + lmbda * off_total).to(torch.float32)
    return loss, dz1, dz2
