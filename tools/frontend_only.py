"""Frontend only (for ncu captures and quick timing on a GPU box).  usage: frontend_only.py [B] [seconds] [iters]"""
import os, sys, types, random
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
from bench import _args_ns, AS_STATS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
L = int(sec * 16000)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
t = torch.arange(L, device=dev, dtype=torch.float32) / 16000.0
wav = 0.1 * torch.randn(B, L, device=dev, generator=g)
for amp in (0.3, 0.1, 0.03):
    f = 100.0 + 6900.0 * torch.rand(B, 1, device=dev, generator=g)
    wav += amp * torch.sin(6.2831853 * f * t[None, :] + 6.2831853 * torch.rand(B, 1, device=dev, generator=g))
wav.clamp_(-1.0, 1.0)
np.random.seed(0); random.seed(0)
fe = S.BatchFrontend(_args_ns(8192), norm_stats=AS_STATS, path="lms", mode="crop")
for _ in range(4):
    fe(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fe(wav)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"B={B} L={L}: {ms*1e3:.1f} us per frontend pass, {B/ms/1e3:.2f} M clips/s, {B*187776/ms/1e6:.0f} GB/s algorithmic")
