"""Multi-GPU Barlow Twins objective on the GLOBAL batch (one process per GPU, torch.distributed / NCCL).

Reference behaviour being replaced: every rank builds its own D x D matrix from its LOCAL batch (local,
unsynchronised BatchNorm) and the ranks `all_reduce(SUM)` those matrices (utils/loss.py:17-21) -- 256 MiB per
rank at D = 8192, and for R > 1 the result is R times the mean of per-rank correlation matrices (SURVEY.md
section 3.3, quirk 1).  North-star semantics implemented here: ONE objective over the global batch of
N_g = sum of local batches, identical (up to rounding) to the single-process reference run on the
rank-ordered concatenation of the batches:

    1. per-dimension statistics of the LOCAL rows (7 numbers per column)            abt_bt_dist_stats_local
       all-gather of those packs (7 D floats per rank)                              NCCL, 229 KB per rank at D = 8192
    2. every rank combines them into the global BatchNorm statistics and writes its
       fp16 STANDARDISED rows into its slot of the gather buffers                   abt_bt_dist_normalize
       all-gather of the standardised embeddings, in place                          NCCL, 2 x N_g x D x 2 B
    3. rank r computes rows [r D/R, (r+1) D/R) of C and of C^T on the tensor cores, its share of the
       off-diagonal loss, and d loss / d z for ALL samples restricted to its dimensions
       (batch-norm backward is per column, so no partial sums cross ranks)          abt_bt_dist_rows_fwd_bwd
    4. all-to-all returns each sample's gradient slice to the rank that owns the sample (NCCL)
    5. a 2-double all-reduce completes the loss

`grad_scale`: DDP averages parameter gradients over ranks, so the local d loss / d z is multiplied by
`world_size` by default; single-process and R-rank runs then produce the same parameter update.

The collective choreography is separated from the compute (`backend`) so that it can be exercised with the gloo
backend on CPUs (tests/test_dist_gloo.py injects a numpy-backed object); the product default is the CUDA
library and nothing else.
"""
from __future__ import annotations

import ctypes as C
import os
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


class CudaBackend:
    """The three compute stages of one step on this rank's GPU, sharing one workspace per (N, R, D) shape."""

    def __init__(self):
        self._ws = {}

    def workspace(self, device, n_local: int, world: int, d: int, row_count: int):
        key = (device.index, n_local, world, d, row_count)
        if key not in self._ws:
            lay = _lib.BtDistLayout()
            _lib.check(_lib.load().abt_bt_dist_layout_query(n_local, world, d, row_count, C.byref(lay)))
            buf = torch.empty(lay.total_bytes + 256, dtype=torch.uint8, device=device)
            base = (buf.data_ptr() + 255) // 256 * 256
            off0 = base - buf.data_ptr()
            ng, pf = n_local * world, int(lay.pack_floats)

            def view(off, numel, dtype, shape):
                nbytes = numel * torch.empty((), dtype=dtype).element_size()
                return buf[off0 + off: off0 + off + nbytes].view(dtype).view(*shape)
            self._ws[key] = dict(buf=buf, base=base, nbytes=int(lay.total_bytes),
                                 zh1=view(lay.zh1, ng * d, torch.float16, (ng, d)), zh2=view(lay.zh2, ng * d, torch.float16, (ng, d)),
                                 pack_local=view(lay.pack_local, pf, torch.float32, (pf,)),
                                 pack_all=view(lay.pack_all, world * pf, torch.float32, (world * pf,)))
        return self._ws[key]

    @staticmethod
    def _stream(dev):
        return torch.cuda.current_stream(dev).cuda_stream

    def stats_local(self, w, z1, z2, world, row_count):
        n, d = int(z1.shape[0]), int(z1.shape[1])
        with torch.cuda.device(z1.device):
            _lib.check(_lib.load().abt_bt_dist_stats_local(z1.data_ptr(), z2.data_ptr(), _DTYPES[z1.dtype], n, world, d, row_count, w["base"],
                                                           self._stream(z1.device)))

    def normalize(self, w, z1, z2, world, rank, row_count, eps, momentum, running_mean, running_var):
        n, d = int(z1.shape[0]), int(z1.shape[1])
        with torch.cuda.device(z1.device):
            _lib.check(_lib.load().abt_bt_dist_normalize(z1.data_ptr(), z2.data_ptr(), _DTYPES[z1.dtype], n, world, rank, d, row_count, float(eps),
                                                         float(momentum), running_mean.data_ptr() if running_mean is not None else None,
                                                         running_var.data_ptr() if running_var is not None else None, w["base"],
                                                         self._stream(z1.device)))

    def rows(self, w, dtype, device, n_local, world, d, row_begin, row_count, alpha, lmbda, hsic, grad_scale, need_mask, phase=0):
        """phase 0: everything; 1: CORR + the dz1 pass (returns parts, dzr1, None); 2: the dz2 pass only (returns None, None, dzr2)."""
        ng = n_local * world
        parts = torch.empty(3, dtype=torch.float64, device=device) if phase != 2 else None
        dzr1 = torch.empty((ng, row_count), dtype=dtype, device=device) if (need_mask & 1) and phase != 2 else None
        dzr2 = torch.empty((ng, row_count), dtype=dtype, device=device) if (need_mask & 2) and phase != 1 else None
        a = _lib.BtDistArgs()
        a.dtype, a.n_local, a.world, a.n_dims = _DTYPES[dtype], n_local, world, d
        a.row_begin, a.row_count = int(row_begin), int(row_count)
        a.alpha, a.lambda_, a.hsic, a.grad_scale = float(alpha), float(lmbda), int(bool(hsic)), float(grad_scale)
        a.need_grad_mask, a.phase = int(need_mask), int(phase)
        a.loss_parts = parts.data_ptr() if parts is not None else w["base"]      # unused in phase 2
        a.dzr1 = dzr1.data_ptr() if dzr1 is not None else None
        a.dzr2 = dzr2.data_ptr() if dzr2 is not None else None
        a.workspace, a.workspace_bytes = w["base"], w["nbytes"]
        with torch.cuda.device(device):
            _lib.check(_lib.load().abt_bt_dist_rows_fwd_bwd(C.byref(a), self._stream(device)))
        return parts, dzr1, dzr2


_CUDA_BACKEND = CudaBackend()


class NativeStep:
    """ONE library call per step (abt_bt_dist_step): private NCCL communicator + communication stream owned by the library.
    The communicator is created once per (process group, device); its NCCL ids (256 bytes) travel over torch.distributed."""

    _comms = {}
    _ws = {}
    _exch = {}          # (device, n, world, d) -> [symmetric buffer, handle, ctypes pointer array, nbytes, epoch] or None (unavailable)

    @classmethod
    def comm(cls, group, device):
        key = (id(group) if group is not None else 0, device.index)
        if key not in cls._comms:
            lib = _lib.load()
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            ident = C.create_string_buffer(256)        # two NCCL ids: one communicator for the large gathers, one for the small exchanges
            if rank == 0:
                _lib.check(lib.abt_comm_unique_id(ident))
            box = [bytes(ident.raw)]
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast_object_list(box, src=src, group=group)
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.abt_comm_create(world, rank, C.create_string_buffer(box[0], 256), C.byref(h)))
            cls._comms[key] = h
        return cls._comms[key]

    @classmethod
    def workspace(cls, device, n, world, d):
        key = (device.index, n, world, d)
        if key not in cls._ws:
            nbytes = C.c_size_t()
            _lib.check(_lib.load().abt_bt_dist_step_workspace_bytes(n, world, d, C.byref(nbytes)))
            cls._ws[key] = (torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device), int(nbytes.value))
        return cls._ws[key]

    @classmethod
    def exchange(cls, group, device, n, world, d):
        """Peer-mapped exchange buffers for the copy-engine embedding gather (torch symmetric memory): every rank allocates one buffer,
        the rendezvous maps all of them into every process.  Returns None when symmetric memory is unavailable (-> NCCL all-gathers) or
        switched off.  ABT_DIST_CE=1 / 0 forces it on / off; by default it is used at two ranks only -- measured (profiles/r2_scaling.md):
        0.908 against 0.923 ms per step at two ranks, but 1.45 against 1.16 ms at eight, where fourteen peer copies per step keep the copy
        engines busier than NCCL's all-gather kernels keep the SMs.  Collective: every rank must call it with the same arguments."""
        key = (id(group) if group is not None else 0, device.index, n, world, d)
        if key in cls._exch:
            return cls._exch[key]
        entry = None
        want = os.environ.get("ABT_DIST_CE")
        if (world == 2 if want is None else want != "0") and world <= 16:
            ok = torch.zeros(1, dtype=torch.int32, device=device)
            try:
                import torch.distributed._symmetric_memory as symm
                nbytes = C.c_size_t()
                _lib.check(_lib.load().abt_bt_dist_exchange_bytes(n, world, d, C.byref(nbytes)))
                pg = group if group is not None else dist.group.WORLD
                buf = symm.empty(nbytes.value, dtype=torch.uint8, device=device)
                hdl = symm.rendezvous(buf, pg.group_name)
                buf.zero_()
                ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
                entry = [buf, hdl, ptrs, int(nbytes.value), 0]
                ok.fill_(1)
            except Exception as e:          # no symmetric memory on this system / build: keep the NCCL gathers
                if dist.get_rank(group) == 0:
                    print(f"ssl_audio_b200: copy-engine exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gathers", flush=True)
            torch.cuda.synchronize(device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # all or nothing; also the barrier after the zero-fill
            if int(ok.item()) == 0:
                entry = None
        cls._exch[key] = entry
        return entry

    @classmethod
    def run(cls, z1, z2, alpha, lmbda, hsic, eps, momentum, running_mean, running_var, need_dz1, need_dz2, grad_scale, group, overlap_hook):
        lib = _lib.load()
        dev = z1.device
        world = dist.get_world_size(group)
        n, d = int(z1.shape[0]), int(z1.shape[1])
        comm = cls.comm(group, dev)
        buf, nbytes = cls.workspace(dev, n, world, d)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = torch.empty_like(z1) if need_dz1 else None
        dz2 = torch.empty_like(z2) if need_dz2 else None
        a = _lib.BtDistStepArgs()
        a.z1, a.z2 = z1.data_ptr(), z2.data_ptr()
        a.dtype, a.n_local, a.n_dims = _DTYPES[z1.dtype], n, d
        a.alpha, a.lambda_, a.hsic = float(alpha), float(lmbda), int(bool(hsic))
        a.eps, a.momentum, a.grad_scale = float(eps), float(momentum), float(grad_scale)
        a.need_grad_mask = (1 if need_dz1 else 0) | (2 if need_dz2 else 0)
        a.loss_out = loss.data_ptr()
        a.dz1 = dz1.data_ptr() if dz1 is not None else None
        a.dz2 = dz2.data_ptr() if dz2 is not None else None
        a.running_mean = running_mean.data_ptr() if running_mean is not None else None
        a.running_var = running_var.data_ptr() if running_var is not None else None
        a.workspace = (buf.data_ptr() + 255) // 256 * 256
        a.workspace_bytes = nbytes
        cb = _lib.OVERLAP_CB(lambda _user: overlap_hook()) if overlap_hook is not None else _lib.OVERLAP_CB()
        a.overlap_cb = cb
        ex = cls.exchange(group, dev, n, world, d)
        if ex is not None:
            ex[4] += 1
            a.exchange_peers = C.cast(ex[2], C.c_void_p)
            a.exchange_bytes = ex[3]
            a.exchange_epoch = ex[4]
        with torch.cuda.device(dev):
            _lib.check(lib.abt_bt_dist_step(C.byref(a), comm, torch.cuda.current_stream(dev).cuda_stream))
        del cb
        return loss, dz1, dz2


def side_stream_hook(side_stream, fn: Callable[[], None], gate: bool = True) -> Callable[[], None]:
    """Build a `comm_overlap_hook` that runs `fn` (kernel launches only -- e.g. `frontend.launch(handle)`) on `side_stream`, gated on the
    objective's stream position: the hook fires on the host as soon as the embedding gathers have been ENQUEUED, which can be a whole step
    before the device gets there; launched ungated, the side-stream kernels fill the SMs early and delay the statistics / standardisation
    kernels the gathers are waiting for (measured at 8 ranks: 1.26 ms per step ungated, 1.06 ms gated, profiles/r2_scaling.md)."""
    ev = torch.cuda.Event()

    def hook():
        if gate:
            ev.record(torch.cuda.current_stream())       # the objective's stream: standardised rows written, gathers launched
            side_stream.wait_event(ev)
        with torch.cuda.stream(side_stream):
            fn()
    return hook


def bt_loss_fwd_bwd_global(z1: torch.Tensor, z2: torch.Tensor, alpha: float, lmbda: float, hsic: bool, *, eps: float = 1e-5,
                           momentum: float = 0.1, running_mean: Optional[torch.Tensor] = None,
                           running_var: Optional[torch.Tensor] = None, need_dz1: bool = True, need_dz2: bool = True,
                           grad_scale: Optional[float] = None, group=None, backend=None, overlap_hook: Optional[Callable[[], None]] = None):
    """Global-batch Barlow Twins loss + local gradients.  z1, z2: this rank's (N, D) embeddings (same N on every rank).
    Returns (loss 0-dim fp32, identical on every rank; dz1 (N, D) or None; dz2 (N, D) or None).

    `overlap_hook`, if given, is called once the all-gather of the standardised embeddings is in flight: whatever it enqueues on the
    current stream (typically the frontend of the NEXT batch) runs while the embeddings cross NVLink."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
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
    if backend is None:
        if not z1.is_cuda:
            raise RuntimeError("embeddings must be CUDA tensors: ssl_audio_b200 has no CPU path")
        if os.environ.get("ABT_DIST_C10D") != "1" and d % world == 0:
            # product default: the whole choreography below as one native call
            return NativeStep.run(z1, z2, alpha, lmbda, hsic, eps, momentum, running_mean, running_var, need_dz1, need_dz2, grad_scale,
                                  group, overlap_hook)
        backend = _CUDA_BACKEND
    w = backend.workspace(z1.device, n, world, d, count)
    # 1. local statistics -> all-gather of the 7 D-float packs
    backend.stats_local(w, z1, z2, world, count)
    dist.all_gather_into_tensor(w["pack_all"], w["pack_local"], group=group)
    # 2. global statistics; standardised local rows into this rank's slot -> in-place all-gather (rank-ordered concatenation =
    #    the single-process global batch)
    backend.normalize(w, z1, z2, world, rank, count, eps, momentum, running_mean, running_var)
    g1 = dist.all_gather_into_tensor(w["zh1"], w["zh1"][rank * n:(rank + 1) * n], group=group, async_op=True)
    g2 = dist.all_gather_into_tensor(w["zh2"], w["zh2"][rank * n:(rank + 1) * n], group=group, async_op=True)
    if overlap_hook is not None:
        overlap_hook()
    g1.wait()
    g2.wait()
    # 3. row block of C / C^T and the gradients of all samples for the dimensions of this rank, in two launches so that
    # 4. the all-to-all returning the dz1 slices to the sample owners ((R, N, D/R) blocks) overlaps the dz2 GEMM
    need_mask = (1 if need_dz1 else 0) | (2 if need_dz2 else 0)

    def exchange_begin(dzr):
        recv = torch.empty((world, n, count), dtype=dzr.dtype, device=dzr.device)
        return recv, dist.all_to_all_single(recv, dzr.view(world, n, count), group=group, async_op=True), dzr

    def exchange_end(pending):
        if pending is None:
            return None
        recv, work, _keep = pending
        work.wait()
        return recv.permute(1, 0, 2).reshape(n, d)        # [src rank = dimension block][n] -> (n, D)

    if need_mask == 3:
        parts, dzr1, _ = backend.rows(w, z1.dtype, z1.device, n, world, d, begin, count, alpha, lmbda, hsic, grad_scale, need_mask, 1)
        p1 = exchange_begin(dzr1)
        _, _, dzr2 = backend.rows(w, z1.dtype, z1.device, n, world, d, begin, count, alpha, lmbda, hsic, grad_scale, need_mask, 2)
        p2 = exchange_begin(dzr2)
    else:
        parts, dzr1, dzr2 = backend.rows(w, z1.dtype, z1.device, n, world, d, begin, count, alpha, lmbda, hsic, grad_scale, need_mask, 0)
        p1 = exchange_begin(dzr1) if dzr1 is not None else None
        p2 = exchange_begin(dzr2) if dzr2 is not None else None
    dz1 = exchange_end(p1)
    dz2 = exchange_end(p2)
    # 5. loss: the off-diagonal partial sums are per row block; the on-diagonal sum is already global and identical on every
    #    rank, so one SUM all-reduce of the three doubles followed by a dot product with constant coefficients finishes the scalar
    dist.all_reduce(parts, group=group)
    key = (parts.device, float(alpha), float(lmbda), bool(hsic), world)
    coef = _COEF.get(key)
    if coef is None:
        coef = torch.tensor([lmbda, 2.0 * lmbda if hsic else 0.0, alpha / world], dtype=torch.float64, device=parts.device)
        _COEF[key] = coef
    loss = torch.dot(parts, coef)
    if hsic:
        loss = loss + lmbda * float(d) * float(d - 1)
    return loss.to(torch.float32), dz1, dz2


_COEF = {}
