"""Loss fwd+bwd only (for ncu captures and quick timing on a GPU box).  usage: loss_only.py [N] [D] [iters]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S

if os.environ.get("ABT_FUSED") is not None:
    from ssl_audio_b200 import _lib as _l
    _l.load().abt_debug_set(9, int(os.environ["ABT_FUSED"]))
if os.environ.get("ABT_FUSED_PDL") is not None:
    from ssl_audio_b200 import _lib as _l
    _l.load().abt_debug_set(10, int(os.environ["ABT_FUSED_PDL"]))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = torch.Generator(device="cuda").manual_seed(1)
z1 = torch.randn(N, D, device="cuda", generator=g)
z2 = (0.6 * z1 + 0.8 * torch.randn(N, D, device="cuda", generator=g)).bfloat16()
z1 = z1.bfloat16()
for _ in range(3):
    S.bt_loss_fwd_bwd(z1, z2, 1.0, 0.005, False)
torch.cuda.synchronize()
import ctypes as C
from ssl_audio_b200 import _lib
lib = _lib.load()
# whole call without per-launch events (programmatic dependent launch stays on)
w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0.record()
for _ in range(iters):
    S.bt_loss_fwd_bwd(z1, z2, 1.0, 0.005, False)
w1.record()
torch.cuda.synchronize()
whole_us = w0.elapsed_time(w1) / iters * 1e3
lib.abt_debug_timing(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss, d1, d2 = S.bt_loss_fwd_bwd(z1, z2, 1.0, 0.005, False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
a, b, c, n = C.c_float(), C.c_float(), C.c_float(), C.c_int()
lib.abt_debug_timing_read(C.byref(a), C.byref(b), C.byref(c), C.byref(n))
lib.abt_debug_timing(0)
if b.value > 1e-4:
    print(f"   stats {a.value*1e3:.1f} us  CORR {b.value*1e3:.1f} us ({2.0*N*D*D/b.value/1e9:.0f} TF)  GRAD {c.value*1e3:.1f} us ({4.0*N*D*D/max(c.value,1e-9)/1e9:.0f} TF)")
else:
    print(f"   stats {a.value*1e3:.1f} us  single-launch kernel {c.value*1e3:.1f} us ({6.0*N*D*D/max(c.value,1e-9)/1e9:.0f} TF algorithmic, {8.0*N*D*D/max(c.value,1e-9)/1e9:.0f} TF executed)")
print(f"N={N} D={D}: {ms*1e3:.1f} us per fwd+bwd with per-launch events, {whole_us:.1f} us without ({6.0*N*D*D/whole_us/1e6:.1f} TFLOP/s algorithmic), loss {float(loss):.4f}")
