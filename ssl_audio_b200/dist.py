"""Multi-GPU evaluation of the Barlow Twins objective on the GLOBAL batch (placeholder until the
row-block kernels land; see DESIGN.md section "Multi-GPU")."""
from __future__ import annotations

import torch


def is_active() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


def bt_loss_fwd_bwd_global(*args, **kwargs):
    raise NotImplementedError("multi-GPU Barlow Twins loss is not implemented yet")
