"""Projector tail fused with the Barlow Twins objective.

Reference: the last bias-free `nn.Linear` of `BarlowTwinsHead` (model.py:22, 25-31) followed by
`BarlowTwinsLoss.forward_loss` (utils/loss.py:15-30).  `proj_tail_forward_loss(loss_module, h_a, h_b, weight)` is

    loss_module.forward_loss(F.linear(h_a, weight), F.linear(h_b, weight))

as ONE autograd node: the Linear of both views runs on the tensor cores (tcgen05, CTA pairs; csrc/bt_loss.cu, LINEAR mode of
`bt_umma_kernel`) and its epilogue, which holds the outputs of both views for the same samples, also produces the per-column
sums the objective's statistics pass would otherwise re-read z for.  The objective continues from that hand-over
(`abt_bt_dist_normalize` / `abt_bt_dist_rows_fwd_bwd` with world = 1); backward turns dz into dh and dW with three plain GEMMs.

Scope: bf16 activations and weight, one weight for both views (Barlow Twins; a BYOL teacher has its own weights), single process.
With N <= 128 rows the one-launch objective (`BarlowTwinsLoss`) is faster than anything this fusion saves; use it there.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import dist as _dist

__all__ = ["proj_tail_forward_loss", "proj_tail_fwd"]

_SCRATCH = {}


def _scratch(dev, n, d):
    key = (dev.index, n, d)
    if key not in _SCRATCH:
        nbytes = C.c_size_t()
        _lib.check(_lib.load().abt_proj_tail_workspace_bytes(n, d, C.byref(nbytes)))
        _SCRATCH[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    return _SCRATCH[key]


def _check(h1, h2, weight):
    for name, t in (("h_a", h1), ("h_b", h2), ("weight", weight)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: ssl_audio_b200 has no CPU path")
        if t.dtype != torch.bfloat16 or t.dim() != 2:
            raise ValueError(f"{name} must be a 2-D bfloat16 tensor (got {t.dtype}, {tuple(t.shape)})")
    if h1.shape != h2.shape or h1.shape[1] != weight.shape[1]:
        raise ValueError(f"shapes do not match: h_a {tuple(h1.shape)}, h_b {tuple(h2.shape)}, weight {tuple(weight.shape)}")


def proj_tail_fwd(h1: torch.Tensor, h2: torch.Tensor, weight: torch.Tensor, pack: torch.Tensor):
    """z1 = h1 W^T, z2 = h2 W^T (bf16) + the 7 x D statistics hand-over written into `pack` (fp32, 7 * D elements)."""
    _check(h1, h2, weight)
    h1, h2, weight = h1.contiguous(), h2.contiguous(), weight.contiguous()
    n, k = int(h1.shape[0]), int(h1.shape[1])
    d = int(weight.shape[0])
    dev = h1.device
    z1 = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    z2 = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    ws = _scratch(dev, n, d)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().abt_proj_tail_fwd(h1.data_ptr(), h2.data_ptr(), weight.data_ptr(), n, k, d, z1.data_ptr(), z2.data_ptr(),
                                                 pack.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return z1, z2


class _TailLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h1, h2, weight, module):
        cfg = module.cfg
        bn = module.bn
        track = bn.training and bn.track_running_stats
        rm = bn.running_mean if track else None
        rv = bn.running_var if track else None
        momentum = bn.momentum if bn.momentum is not None else 0.1
        h1d, h2d, wd = h1.detach().contiguous(), h2.detach().contiguous(), weight.detach().contiguous()
        n, d = int(h1d.shape[0]), int(wd.shape[0])
        dev = h1d.device
        need1 = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        need2 = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        be = _dist._CUDA_BACKEND
        w = be.workspace(dev, n, 1, d, d)
        z1, z2 = proj_tail_fwd(h1d, h2d, wd, w["pack_all"])
        be.normalize(w, z1, z2, 1, 0, d, bn.eps, momentum, rm, rv)
        mask = (1 if need1 else 0) | (2 if need2 else 0)
        parts, dz1, dz2 = be.rows(w, torch.bfloat16, dev, n, 1, d, 0, d, cfg.alpha, cfg.lmbda, cfg.HSIC, 1.0, mask, 0)
        key = (parts.device, float(cfg.alpha), float(cfg.lmbda), bool(cfg.HSIC), 1)
        coef = _dist._COEF.get(key)
        if coef is None:
            coef = torch.tensor([cfg.lmbda, 2.0 * cfg.lmbda if cfg.HSIC else 0.0, cfg.alpha], dtype=torch.float64, device=parts.device)
            _dist._COEF[key] = coef
        loss = torch.dot(parts, coef)
        if cfg.HSIC:
            loss = loss + cfg.lmbda * float(d) * float(d - 1)
        if track:
            module._pending_batches += 2
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(h1d, h2d, wd, dz1 if dz1 is not None else empty, dz2 if dz2 is not None else empty)
        return loss.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("proj_tail_forward_loss: the gradients of this evaluation were already consumed (they are scaled in place by "
                               "grad_output); call it again instead of back-propagating twice through the same graph")
        ctx.consumed = True
        h1, h2, w, dz1, dz2 = ctx.saved_tensors
        g1 = dz1 if dz1.numel() else None
        g2 = dz2 if dz2.numel() else None
        dh1 = dh2 = dw = None
        # dz = d loss / d z comes straight from the GRAD launch (this call's own buffers): scaled by grad_output in place with one launch
        # (a no-op when it is 1), then the Linear's own backward is three library GEMMs, exactly what nn.Linear's backward runs
        ref = g1 if g1 is not None else g2
        if ref is not None:
            scale = grad_out.detach().to(torch.float32).contiguous()
            with torch.cuda.device(ref.device):
                _lib.check(_lib.load().abt_scale_inplace(g1.data_ptr() if g1 is not None else None, g2.data_ptr() if g2 is not None else None,
                                                         ref.numel(), _lib.DTYPE_BF16, scale.data_ptr(),
                                                         torch.cuda.current_stream(ref.device).cuda_stream))
        if ctx.needs_input_grad[0]:
            dh1 = g1 @ w
        if ctx.needs_input_grad[1]:
            dh2 = g2 @ w
        if ctx.needs_input_grad[2]:
            dw = torch.cat((g1, g2)).t() @ torch.cat((h1, h2))       # one GEMM over both views: a single rounding of dW
        return dh1, dh2, dw, None


def proj_tail_forward_loss(loss_module, h_a: torch.Tensor, h_b: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """`loss_module.forward_loss(F.linear(h_a, weight), F.linear(h_b, weight))` in one autograd node (see the module docstring).
    `loss_module` is a `ssl_audio_b200.BarlowTwinsLoss` (its cfg, BatchNorm buffers and counters are used and updated)."""
    _check(h_a, h_b, weight)
    if weight.shape[0] != loss_module.cfg.projector_out_dim:
        raise ValueError(f"expected a weight with {loss_module.cfg.projector_out_dim} output features, got {weight.shape[0]}")
    if _dist.is_active():
        raise NotImplementedError("the fused projector tail is single-process; under torch.distributed use the head's Linear + BarlowTwinsLoss")
    if not loss_module.bn.training and loss_module.bn.track_running_stats:
        raise NotImplementedError("eval-mode BatchNorm is not part of the accelerated path (see BarlowTwinsLoss.forward_loss)")
    return _TailLossFn.apply(h_a, h_b, weight, loss_module)
