"""ctypes binding of libabt_b200.so (include/abt_b200.h).

There is NO fallback: if the library is missing and cannot be built, or a compute call is made
on a non-sm_100 device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABT_LIB") or os.path.join(_HERE, "libabt_b200.so")     # ABT_LIB: A/B timing of two builds on one box

ABT_ERR_ARG = -1

DTYPE_BF16, DTYPE_F16, DTYPE_F32 = 0, 1, 2


class MelConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("n_mels", C.c_int32), ("f_min", C.c_float), ("f_max", C.c_float), ("apply_norm", C.c_int32),
        ("norm_mean", C.c_float), ("norm_std", C.c_float),
    ]


class ViewParams(C.Structure):
    _fields_ = [
        ("z_kind", C.c_int32), ("z_index", C.c_int32), ("w_x", C.c_float), ("w_z", C.c_float),
        ("i", C.c_int32), ("j", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("head", C.c_float), ("tail", C.c_float), ("flags", C.c_int32), ("out_index", C.c_int32),
        ("g_lambda", C.c_float), ("g_keep", C.c_float), ("reserved", C.c_int32 * 2),
    ]


class ViewsArgs(C.Structure):
    _fields_ = [
        ("n_clips", C.c_int32), ("n_views", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("canvas_h", C.c_int32), ("canvas_w", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("param_stride", C.c_int32), ("view_offset", C.c_int32),
        ("x", C.c_void_p), ("x_slot", C.c_void_p), ("x_slot_stride", C.c_int64),
        ("bank", C.c_void_p), ("bank_slot_stride", C.c_int64), ("params", C.c_void_p),
        ("outs", C.c_void_p * 8), ("noise", C.c_void_p), ("noise_views", C.c_int32),
    ]


class PlanConfig(C.Structure):
    _fields_ = [
        ("mixup", C.c_int32), ("rrc", C.c_int32), ("rlf", C.c_int32), ("mixup_ratio_d", C.c_double),
        ("n_memory", C.c_int32), ("ring_slots", C.c_int32), ("n_global", C.c_int32),
        ("in_h", C.c_int32), ("in_w", C.c_int32), ("canvas_h", C.c_int32), ("canvas_w", C.c_int32),
        ("freq_scale", C.c_double * 2), ("time_scale", C.c_double * 2),
        ("n_local", C.c_int32), ("local_h", C.c_int32), ("local_w", C.c_int32),
        ("local_scale", C.c_double * 2), ("fader_gain", C.c_double),
        ("gnoise", C.c_int32), ("gnoise_ratio_d", C.c_double),
    ]


class BtArgs(C.Structure):
    _fields_ = [
        ("z1", C.c_void_p), ("z2", C.c_void_p), ("dtype", C.c_int32), ("n_rows", C.c_int32), ("n_dims", C.c_int32),
        ("alpha", C.c_float), ("lambda_", C.c_float), ("hsic", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
        ("grad_scale", C.c_float), ("need_grad_mask", C.c_int32), ("loss_out", C.c_void_p), ("dz1", C.c_void_p),
        ("dz2", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


class BtRowsArgs(C.Structure):
    _fields_ = [
        ("zg1", C.c_void_p), ("zg2", C.c_void_p), ("dtype", C.c_int32), ("n_rows", C.c_int32), ("n_dims", C.c_int32),
        ("row_begin", C.c_int32), ("row_count", C.c_int32), ("alpha", C.c_float), ("lambda_", C.c_float), ("hsic", C.c_int32),
        ("eps", C.c_float), ("momentum", C.c_float), ("grad_scale", C.c_float), ("need_grad_mask", C.c_int32),
        ("loss_parts", C.c_void_p), ("dzr1", C.c_void_p), ("dzr2", C.c_void_p), ("running_mean", C.c_void_p),
        ("running_var", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class BtDistLayout(C.Structure):
    _fields_ = [("total_bytes", C.c_size_t), ("zh1", C.c_size_t), ("zh2", C.c_size_t), ("pack_local", C.c_size_t),
                ("pack_all", C.c_size_t), ("pack_floats", C.c_size_t)]


class BtDistArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("n_local", C.c_int32), ("world", C.c_int32), ("n_dims", C.c_int32), ("row_begin", C.c_int32),
        ("row_count", C.c_int32), ("alpha", C.c_float), ("lambda_", C.c_float), ("hsic", C.c_int32), ("grad_scale", C.c_float),
        ("need_grad_mask", C.c_int32), ("phase", C.c_int32), ("loss_parts", C.c_void_p), ("dzr1", C.c_void_p), ("dzr2", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class OptTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("aux", C.c_void_p), ("n", C.c_longlong), ("flags", C.c_int32), ("chunk0", C.c_int32)]


OVERLAP_CB = C.CFUNCTYPE(None, C.c_void_p)


class BtDistStepArgs(C.Structure):
    _fields_ = [
        ("z1", C.c_void_p), ("z2", C.c_void_p), ("dtype", C.c_int32), ("n_local", C.c_int32), ("n_dims", C.c_int32),
        ("alpha", C.c_float), ("lambda_", C.c_float), ("hsic", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
        ("grad_scale", C.c_float), ("need_grad_mask", C.c_int32), ("loss_out", C.c_void_p), ("dz1", C.c_void_p), ("dz2", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("overlap_cb", OVERLAP_CB), ("overlap_user", C.c_void_p),
        ("exchange_peers", C.c_void_p), ("exchange_bytes", C.c_size_t), ("exchange_epoch", C.c_uint32),
    ]


# name -> (restype, argtypes); must list every symbol of include/abt_b200.h
SIGNATURES = {
    "abt_version": (C.c_int, []),
    "abt_last_error": (C.c_char_p, []),
    "abt_device_check": (C.c_int, []),
    "abt_logmel_plan_create": (C.c_int, [C.POINTER(MelConfig), C.POINTER(C.c_void_p)]),
    "abt_logmel_plan_destroy": (C.c_int, [C.c_void_p]),
    "abt_logmel_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "abt_logmel_crop_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "abt_wav_span_len": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "abt_wav_span_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "abt_logmel_span_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p]),
    "abt_lms_crop_norm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "abt_views_fwd": (C.c_int, [C.POINTER(ViewsArgs), C.c_void_p]),
    "abt_bank_push": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "abt_normalize_batch_workspace_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "abt_normalize_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abt_mean_std": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abt_running_norm_workspace_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "abt_running_norm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abt_planner_create": (C.c_int, [C.POINTER(PlanConfig), C.POINTER(C.c_void_p)]),
    "abt_planner_destroy": (C.c_int, [C.c_void_p]),
    "abt_planner_set_numpy_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "abt_planner_get_numpy_state": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "abt_planner_set_pyrandom_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "abt_planner_get_pyrandom_state": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "abt_planner_bank_len": (C.c_int, [C.c_void_p]),
    "abt_planner_bank_reset": (C.c_int, [C.c_void_p]),
    "abt_planner_plan_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abt_planner_packed_bytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                           C.POINTER(C.c_size_t)]),
    "abt_planner_plan_batch_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "abt_planner_plan_batch_global": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "abt_bt_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "abt_bt_loss_fwd_bwd": (C.c_int, [C.POINTER(BtArgs), C.c_void_p]),
    "abt_scale_inplace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "abt_bt_rows_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "abt_bt_loss_rows_fwd_bwd": (C.c_int, [C.POINTER(BtRowsArgs), C.c_void_p]),
    "abt_bt_dist_layout_query": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(BtDistLayout)]),
    "abt_bt_dist_stats_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "abt_bt_dist_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abt_bt_dist_rows_fwd_bwd": (C.c_int, [C.POINTER(BtDistArgs), C.c_void_p]),
    "abt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "abt_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "abt_comm_destroy": (C.c_int, [C.c_void_p]),
    "abt_bt_dist_step_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "abt_bt_dist_exchange_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "abt_bt_dist_step": (C.c_int, [C.POINTER(BtDistStepArgs), C.c_void_p, C.c_void_p]),
    "abt_opt_chunk_elems": (C.c_int, []),
    "abt_lars_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "abt_ema_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "abt_set_reserved_sms": (C.c_int, [C.c_int]),
    "abt_proj_tail_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "abt_proj_tail_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "abt_debug_set": (C.c_int, [C.c_int, C.c_int]),
    "abt_debug_launch_count": (C.c_longlong, [C.c_int]),
    "abt_debug_timing": (C.c_int, [C.c_int]),
    "abt_debug_timing_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "abt_debug_ws_offsets": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load (building in-tree first if needed) libabt_b200.so.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        # debugging aid: ABT_CTA_GROUP=1 selects the single-CTA tensor-core kernel instead of the CTA-pair one
        if os.environ.get("ABT_CTA_GROUP") in ("1", "2"):
            lib.abt_debug_set(6, int(os.environ["ABT_CTA_GROUP"]))
        if os.environ.get("ABT_DIST_XCHG") in ("0", "1"):       # 0: multi-GPU step without the exchange schedule
            lib.abt_debug_set(7, int(os.environ["ABT_DIST_XCHG"]))
        if os.environ.get("ABT_COMM_MAX_CTAS") is not None:     # CTA cap of the private NCCL communicators (0 = NCCL's default)
            lib.abt_debug_set(13, int(os.environ["ABT_COMM_MAX_CTAS"]))
        if os.environ.get("ABT_DIST_RESERVE_SMS") is not None:  # SMs the tensor-core kernels leave to NCCL inside the multi-GPU step
            lib.abt_debug_set(14, int(os.environ["ABT_DIST_RESERVE_SMS"]))
        _lib = lib
    return _lib


def last_error() -> str:
    return load().abt_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Translate an abt_status into the Python exception the reference's callers would see."""
    if rc == 0:
        return
    msg = last_error()
    if rc == ABT_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(f"libabt_b200 error {rc}: {msg}")
