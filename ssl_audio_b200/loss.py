"""GPU mirror of the reference's Barlow Twins objective (reference: utils/loss.py:8-48, utils/utils.py:23-27).

`BarlowTwinsLoss(cfg, ncrops)` keeps the reference's constructor, `forward_loss(z1, z2)` and
`forward(student_output, teacher_output, ngcrops_each=1)`, reads the same cfg fields
(projector_out_dim, HSIC, alpha, lmbda) and exposes the same `state_dict` keys
(`bn.running_mean`, `bn.running_var`, `bn.num_batches_tracked`).  Loss AND gradients are produced by
one call into libabt_b200 (csrc/bt_loss.cu) during forward; backward only scales the stored
gradients by `grad_output`, which also covers GradScaler.  There is no PyTorch fallback.

Multi-GPU: when torch.distributed is initialised with world_size > 1 the objective is evaluated on
the GLOBAL batch (global batch-norm statistics, C over all N_g rows) -- see ssl_audio_b200/dist.py
and DESIGN.md for how this differs from the reference's unsynchronised local-BN x all_reduce(SUM).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib

__all__ = ["BarlowTwinsLoss", "off_diagonal", "bt_loss_fwd_bwd"]

_DTYPES = {torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}

# fp16 embeddings (the reference's AMP mode is fp16 autocast + GradScaler, main.py:84,137): the library writes the gradients in
# the input dtype during forward, BEFORE autograd hands over the loss scale, so small d loss / d z (< 6e-5) would be fp16
# subnormals.  They are therefore stored pre-multiplied by a power of two and backward divides it back out of grad_output.
_FP16_PRESCALE = 4096.0


def off_diagonal(x: torch.Tensor) -> torch.Tensor:
    """Flattened view of the off-diagonal elements of a square matrix (reference: utils/utils.py:23-27)."""
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()


class _Workspace:
    """Per-(device, N, D, dtype) scratch kept alive between steps (H is D*D bf16: 128 MiB at D = 8192)."""

    def __init__(self):
        self._buf = {}

    def get(self, device: torch.device, n: int, d: int, dtype_code: int) -> Tuple[torch.Tensor, int]:
        key = (device.index, n, d, dtype_code)
        if key not in self._buf:
            nbytes = C.c_size_t()
            _lib.check(_lib.load().abt_bt_workspace_bytes(n, d, dtype_code, C.byref(nbytes)))
            self._buf[key] = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
        buf = self._buf[key]
        ptr = (buf.data_ptr() + 255) // 256 * 256
        return buf, ptr


_WS = _Workspace()


def bt_loss_fwd_bwd(z1: torch.Tensor, z2: torch.Tensor, alpha: float, lmbda: float, hsic: bool, *, eps: float = 1e-5,
                    momentum: float = 0.1, running_mean: Optional[torch.Tensor] = None, running_var: Optional[torch.Tensor] = None,
                    need_dz1: bool = True, need_dz2: bool = True, grad_scale: float = 1.0):
    """One fused evaluation: returns (loss 0-dim fp32, dz1 or None, dz2 or None).  z1, z2: (N, D) CUDA, same dtype."""
    if not (z1.is_cuda and z2.is_cuda):
        raise RuntimeError("embeddings must be CUDA tensors: ssl_audio_b200 has no CPU path")
    if z1.dim() != 2 or z1.shape != z2.shape:
        raise ValueError(f"z1 and z2 must both be (N, D); got {tuple(z1.shape)} and {tuple(z2.shape)}")
    if z1.dtype != z2.dtype or z1.dtype not in _DTYPES:
        raise ValueError(f"unsupported embedding dtypes {z1.dtype}, {z2.dtype}")
    lib = _lib.load()
    z1 = z1.contiguous()
    z2 = z2.contiguous()
    n, d = int(z1.shape[0]), int(z1.shape[1])
    code = _DTYPES[z1.dtype]
    dev = z1.device
    with torch.cuda.device(dev):
        buf, ws_ptr = _WS.get(dev, n, d, code)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = torch.empty_like(z1) if need_dz1 else None
        dz2 = torch.empty_like(z2) if need_dz2 else None
        a = _lib.BtArgs()
        a.z1, a.z2 = z1.data_ptr(), z2.data_ptr()
        a.dtype, a.n_rows, a.n_dims = code, n, d
        a.alpha, a.lambda_, a.hsic = float(alpha), float(lmbda), int(bool(hsic))
        a.eps, a.momentum, a.grad_scale = float(eps), float(momentum), float(grad_scale)
        a.need_grad_mask = (1 if need_dz1 else 0) | (2 if need_dz2 else 0)
        a.loss_out = loss.data_ptr()
        a.dz1 = dz1.data_ptr() if dz1 is not None else None
        a.dz2 = dz2.data_ptr() if dz2 is not None else None
        a.running_mean = running_mean.data_ptr() if running_mean is not None else None
        a.running_var = running_var.data_ptr() if running_var is not None else None
        a.workspace, a.workspace_bytes = ws_ptr, buf.numel() - 256
        _lib.check(lib.abt_bt_loss_fwd_bwd(C.byref(a), torch.cuda.current_stream(dev).cuda_stream))
    return loss, dz1, dz2


class _BTLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, module):
        cfg = module.cfg
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        bn = module.bn
        track = bn.training and bn.track_running_stats
        rm = bn.running_mean if track else None
        rv = bn.running_var if track else None
        if rm is not None and (rm.device != z1.device or rm.dtype != torch.float32):
            raise RuntimeError("BarlowTwinsLoss buffers must be fp32 on the embeddings' device (call .cuda() as main.py:422 does)")
        if track and bn.momentum is None:
            raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative moving average) is not supported; the reference uses the default 0.1")
        momentum = bn.momentum if bn.momentum is not None else 0.1
        pre = _FP16_PRESCALE if z1.dtype == torch.float16 else 1.0
        from . import dist as _dist
        if _dist.is_active():
            gs = module.grad_scale if module.grad_scale is not None else float(torch.distributed.get_world_size())
            loss, dz1, dz2 = _dist.bt_loss_fwd_bwd_global(z1.detach(), z2.detach(), cfg.alpha, cfg.lmbda, cfg.HSIC, eps=bn.eps,
                                                          momentum=momentum, running_mean=rm, running_var=rv, need_dz1=need1, need_dz2=need2,
                                                          grad_scale=gs * pre, overlap_hook=module.comm_overlap_hook)
        else:
            loss, dz1, dz2 = bt_loss_fwd_bwd(z1.detach(), z2.detach(), cfg.alpha, cfg.lmbda, cfg.HSIC, eps=bn.eps, momentum=momentum,
                                             running_mean=rm, running_var=rv, need_dz1=need1, need_dz2=need2, grad_scale=pre)
        ctx.prescale = pre
        if track:
            module._pending_batches += 2    # BatchNorm is applied to z1 and then to z2 (utils/loss.py:17); flushed lazily
        ctx.save_for_backward(dz1 if dz1 is not None else torch.empty(0, device=z1.device),
                              dz2 if dz2 is not None else torch.empty(0, device=z1.device))
        ctx.has = (dz1 is not None, dz2 is not None)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("BarlowTwinsLoss: the gradients of this evaluation were already consumed (they are scaled in place by "
                               "grad_output); call the loss again instead of back-propagating twice through the same graph")
        ctx.consumed = True
        dz1, dz2 = ctx.saved_tensors
        # the stored gradients are this call's own buffers (never exposed before): scale both in place with one launch
        g1 = dz1 if ctx.has[0] else None
        g2 = dz2 if ctx.has[1] else None
        ref = g1 if g1 is not None else g2
        if ref is not None:
            scale = grad_out.detach().to(torch.float32).contiguous()
            if ctx.prescale != 1.0:
                scale = scale / ctx.prescale
            with torch.cuda.device(ref.device):
                _lib.check(_lib.load().abt_scale_inplace(g1.data_ptr() if g1 is not None else None, g2.data_ptr() if g2 is not None else None,
                                                         ref.numel(), _DTYPES[ref.dtype], scale.data_ptr(),
                                                         torch.cuda.current_stream(ref.device).cuda_stream))
        return g1, g2, None


class BarlowTwinsLoss(nn.Module):
    def __init__(self, cfg, ncrops):
        super().__init__()
        self.cfg = cfg
        self.ncrops = ncrops
        # holds running_mean / running_var / num_batches_tracked exactly like the reference module (utils/loss.py:13);
        # the normalisation itself happens inside the CUDA kernels
        self.bn = nn.BatchNorm1d(cfg.projector_out_dim, affine=False)
        # multi-GPU only: local gradients are multiplied by this; None = world_size, which cancels DDP's gradient
        # averaging so that R-rank and single-process runs give the same parameter update
        self.grad_scale = None
        # multi-GPU only: optional callable invoked while the embedding all-gather is in flight (e.g. the next batch's frontend)
        self.comm_overlap_hook = None
        # `bn.num_batches_tracked` is bookkeeping only (momentum is fixed): count on the host and fold the count into the
        # buffer when somebody looks at it (state_dict / explicit flush) instead of launching a kernel every step
        self._pending_batches = 0
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: module.flush_counters())
        # a checkpoint's counter replaces (not adds to) the steps taken before it was loaded
        self.register_load_state_dict_pre_hook(lambda module, *a, **k: setattr(module, "_pending_batches", 0))

    def flush_counters(self) -> None:
        if self._pending_batches:
            self.bn.num_batches_tracked += self._pending_batches
            self._pending_batches = 0

    def forward_loss(self, z1, z2):
        if z1.shape[-1] != self.cfg.projector_out_dim:
            raise ValueError(f"expected embeddings with {self.cfg.projector_out_dim} dims, got {z1.shape[-1]}")
        if not self.bn.training and self.bn.track_running_stats:
            raise NotImplementedError("eval-mode BatchNorm (normalising with the running statistics) is not part of the accelerated path; "
                                      "the reference never switches the loss module to eval (main.py:84-139 trains only)")
        return _BTLossFn.apply(z1, z2, self)

    def forward(self, student_output, teacher_output, ngcrops_each=1):
        # pairing loop of the reference (utils/loss.py:32-48)
        student_out = student_output.chunk(self.ncrops - (2 - ngcrops_each))
        teacher_out = teacher_output.chunk(ngcrops_each)
        total_loss = None
        n_loss_terms = 0
        for q in range(len(teacher_out)):
            for v in range(len(student_out)):
                if len(teacher_out) > 1 and q == v:
                    continue
                term = self.forward_loss(teacher_out[q], student_out[v])
                total_loss = term if total_loss is None else total_loss + term
                n_loss_terms += 1
        if n_loss_terms > 1:                      # (0 + x) / 1 == x: no arithmetic (and no launches) for the single-term case
            total_loss = total_loss / n_loss_terms
        return total_loss
