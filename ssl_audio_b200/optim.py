"""Step-adjacent optimiser math on the GPU path (reference: utils/utils.py:150-189 `LARS`, utils/utils.py:311-331 `EMA` /
`update_moving_average`).

Same class names, constructor arguments and update rules as the reference; the per-parameter Python loops are replaced by
multi-tensor kernels of libabt_b200 (csrc/optim.cu): one device table describes every tensor of a step, so `LARS.step()` costs two
launches and `update_moving_average` one, whatever the number of parameters (a ResNet-18 has 62 tensors, ViT-Base 150+: the
reference issues ~10 launches and two `torch.norm` per tensor).  fp32 CUDA parameters only; there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib

__all__ = ["LARS", "EMA", "update_moving_average", "exclude_bias_and_norm"]


def exclude_bias_and_norm(p) -> bool:
    """utils/utils.py:158-159."""
    return p.ndim == 1


class _Table:
    """Device table (abt_opt_tensor records + chunk map) for a list of tensors; rebuilt only when a pointer changes."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.n_tensors = self.n_chunks = 0
        self.chunk = int(_lib.load().abt_opt_chunk_elems())

    def update(self, device, ps: Sequence[torch.Tensor], gs: Sequence[torch.Tensor], auxs: Sequence, flags: Sequence[int]):
        key = tuple((p.data_ptr(), g.data_ptr(), a.data_ptr() if a is not None else 0, p.numel(), f) for p, g, a, f in zip(ps, gs, auxs, flags))
        if key == self.key:
            return
        n = len(ps)
        recs = (_lib.OptTensor * n)()
        chunk_tensor: List[int] = []
        for t, (p, g, a, f) in enumerate(zip(ps, gs, auxs, flags)):
            r = recs[t]
            r.p, r.g, r.aux = p.data_ptr(), g.data_ptr(), (a.data_ptr() if a is not None else None)
            r.n, r.flags, r.chunk0 = p.numel(), int(f), len(chunk_tensor)
            chunk_tensor.extend([t] * ((p.numel() + self.chunk - 1) // self.chunk))
        raw = bytes(recs)
        rec_bytes = len(raw)
        host = torch.empty(rec_bytes + 4 * len(chunk_tensor), dtype=torch.uint8).pin_memory()
        host[:rec_bytes] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        host[rec_bytes:] = torch.tensor(chunk_tensor, dtype=torch.int32).view(torch.uint8)
        self.dev = host.to(device, non_blocking=True)
        self._host = host                                  # keep the pinned staging buffer alive until the copy has run
        self.rec_bytes = rec_bytes
        self.n_tensors, self.n_chunks = n, len(chunk_tensor)
        self.partial = torch.empty(2 * self.n_chunks + self.n_tensors, dtype=torch.float32, device=device)     # chunk norms + one trust ratio per tensor
        self.key = key

    @property
    def recs_ptr(self) -> int:
        return self.dev.data_ptr()

    @property
    def chunks_ptr(self) -> int:
        return self.dev.data_ptr() + self.rec_bytes


def _check(ts: Sequence[torch.Tensor], what: str):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError(f"{what} must be CUDA tensors: ssl_audio_b200 has no CPU path")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{what} must be contiguous float32 tensors")


class LARS(torch.optim.Optimizer):
    """LARS with the reference's constructor and update rule (utils/utils.py:150-189)."""

    def __init__(self, params, lr, weight_decay=0, momentum=0.9, eta=0.001, weight_decay_filter=False, lars_adaptation_filter=False):
        defaults = dict(lr=lr, weight_decay=weight_decay, momentum=momentum, eta=eta, weight_decay_filter=weight_decay_filter,
                        lars_adaptation_filter=lars_adaptation_filter)
        super().__init__(params, defaults)
        self._tables = {}

    def exclude_bias_and_norm(self, p):
        return exclude_bias_and_norm(p)

    @torch.no_grad()
    def step(self):
        lib = _lib.load()
        for gi, g in enumerate(self.param_groups):
            ps, gs, mus, flags = [], [], [], []
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "mu" not in st:
                    st["mu"] = torch.zeros_like(p)
                excl = self.exclude_bias_and_norm(p)
                f = (1 if (not g["weight_decay_filter"] or not excl) else 0) | (2 if (not g["lars_adaptation_filter"] or not excl) else 0)
                ps.append(p.data); gs.append(p.grad.data); mus.append(st["mu"]); flags.append(f)
            if not ps:
                continue
            _check(ps, "parameters"); _check(gs, "gradients")
            dev = ps[0].device
            tab = self._tables.setdefault((gi, dev.index), _Table())
            with torch.cuda.device(dev):
                tab.update(dev, ps, gs, mus, flags)
                _lib.check(lib.abt_lars_step(tab.recs_ptr, tab.chunks_ptr, tab.n_tensors, tab.n_chunks, float(g["lr"]), float(g["weight_decay"]),
                                             float(g["momentum"]), float(g["eta"]), tab.partial.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))


class EMA:
    """utils/utils.py:311-325."""

    def __init__(self, beta):
        super().__init__()
        self.beta = beta

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new


_EMA_TABLES = {}


@torch.no_grad()
def update_moving_average(ema_updater, ma_model, current_model):
    """utils/utils.py:328-331: ma = beta * ma + (1 - beta) * current for every parameter pair, in ONE launch (in place on ma)."""
    cur = [p.data for p in current_model.parameters()]
    ma = [p.data for p in ma_model.parameters()]
    if len(cur) != len(ma):
        raise ValueError("the two models have different parameter lists")
    if not ma:
        return
    _check(ma, "moving-average parameters"); _check(cur, "online parameters")
    dev = ma[0].device
    tab = _EMA_TABLES.setdefault((id(ma_model), id(current_model), dev.index), _Table())
    with torch.cuda.device(dev):
        tab.update(dev, ma, cur, [None] * len(ma), [0] * len(ma))
        _lib.check(_lib.load().abt_ema_update(tab.recs_ptr, tab.chunks_ptr, tab.n_tensors, tab.n_chunks, float(ema_updater.beta),
                                              torch.cuda.current_stream(dev).cuda_stream))
