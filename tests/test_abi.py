"""The C-ABI library loads on a machine without a GPU and exports every symbol include/abt_b200.h declares;
argument validation and the 'no fallback' rule work without touching a device."""
import ctypes as C
import os
import re

import pytest

from ssl_audio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "abt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 23
    for n in names:
        assert hasattr(lib, n), f"{n} declared in abt_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in ssl_audio_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.abt_version() == 100
    assert isinstance(lib.abt_last_error(), bytes)


def test_struct_sizes_match_header(tmp_path):
    """sizeof of every struct that crosses the boundary, as gcc lays it out from include/abt_b200.h, equals the ctypes mirror."""
    import subprocess
    pairs = {"abt_view_params": _lib.ViewParams, "abt_mel_config": _lib.MelConfig, "abt_bt_args": _lib.BtArgs,
             "abt_views_args": _lib.ViewsArgs, "abt_plan_config": _lib.PlanConfig, "abt_bt_rows_args": _lib.BtRowsArgs,
             "abt_bt_dist_layout": _lib.BtDistLayout, "abt_bt_dist_args": _lib.BtDistArgs, "abt_bt_dist_step_args": _lib.BtDistStepArgs,
             "abt_opt_tensor": _lib.OptTensor}
    src = tmp_path / "sz.c"
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "abt_b200.h")
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in pairs)
    src.write_text(f'#include <stdio.h>\n#include "{header}"\nint main(void){{{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = dict(zip(out[::2], (int(v) for v in out[1::2])))
    for name, cls in pairs.items():
        assert sizes[name] == C.sizeof(cls), name
    assert sizes["abt_view_params"] == 64


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    nbytes = C.c_size_t()
    assert lib.abt_bt_workspace_bytes(1024, 8192, 0, C.byref(nbytes)) == 0
    assert nbytes.value > 2 * 8192 * 8192                      # the fp16 correlation matrix alone is 128 MiB
    assert lib.abt_bt_workspace_bytes(128, 8192, 0, C.byref(nbytes)) == 0
    assert nbytes.value < 32 << 20                              # N <= 128: the one-launch kernel keeps C on chip
    assert lib.abt_bt_workspace_bytes(128, 100, 0, C.byref(nbytes)) == _lib.ABT_ERR_ARG
    assert b"multiple of 64" in lib.abt_last_error()
    assert lib.abt_bt_workspace_bytes(1, 128, 0, C.byref(nbytes)) == _lib.ABT_ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(lib.abt_bt_loss_fwd_bwd(None, None))
    cfg = _lib.MelConfig(16000, 512, 512, 160, 64, 60.0, 7800.0, 0, 0.0, 1.0)
    h = C.c_void_p()
    assert lib.abt_logmel_plan_create(C.byref(cfg), C.byref(h)) == _lib.ABT_ERR_ARG   # n_fft must be 1024
    assert lib.abt_running_norm_workspace_bytes(-1, C.byref(nbytes)) == _lib.ABT_ERR_ARG
    assert lib.abt_running_norm_workspace_bytes(8, C.byref(nbytes)) == 0 and nbytes.value >= 8 * 24
    assert lib.abt_running_norm(None, 4, 6144, 10, None, None, None, None) == _lib.ABT_ERR_ARG
    va = _lib.ViewsArgs()
    va.n_clips, va.n_views, va.in_h, va.in_w, va.canvas_h, va.canvas_w, va.out_h, va.out_w = 2, 2, 64, 96, 64, 144, 64, 96
    va.x, va.params, va.noise, va.noise_views = 1 << 20, 1 << 20, 1 << 20, 1        # (never dereferenced: validation fails first)
    assert lib.abt_views_fwd(C.byref(va), None) == _lib.ABT_ERR_ARG
    assert b"noise" in lib.abt_last_error()


def test_no_cpu_fallback():
    import torch
    import ssl_audio_b200 as S
    if torch.cuda.is_available():
        pytest.skip("checks the behaviour of a GPU-less host")
    assert _lib.load().abt_device_check() != 0
    import types
    crit = S.BarlowTwinsLoss(types.SimpleNamespace(projector_out_dim=64, HSIC=False, alpha=1.0, lmbda=0.005), ncrops=2)
    with pytest.raises(RuntimeError):
        crit.forward_loss(torch.zeros(8, 64), torch.zeros(8, 64))
    with pytest.raises(RuntimeError):
        S.LogMelSpectrogram(16000, 1024, 1024, 160, 64, 60, 7800)(torch.zeros(1, 4000))
    with pytest.raises(RuntimeError):
        S.RandomResizeCrop()(torch.zeros(1, 64, 96))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ssl_audio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # no import / attribute use / path of the checker anywhere in the product (prose mentions are fine)
                assert not re.search(r"(^|\n)\s*(from|import)\s+oracle\b|\boracle\.\w|oracle/|import_module\([^)]*oracle", src), f
