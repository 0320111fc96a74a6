"""Projector tail (f2) alone: python tools/proj_only.py [N K D iters]  -> device time of abt_proj_tail_fwd (both launches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssl_audio_b200.projector import proj_tail_fwd
n, k, d, it = (int(v) for v in (sys.argv[1:5] + ["1024", "8192", "8192", "20"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
h1 = torch.relu(torch.randn(n, k, device=dev, generator=g)).bfloat16()
h2 = torch.relu(torch.randn(n, k, device=dev, generator=g)).bfloat16()
w = (torch.randn(d, k, device=dev, generator=g) / k ** 0.5).bfloat16()
pack = torch.zeros(7 * d, device=dev)
for _ in range(3):
    proj_tail_fwd(h1, h2, w, pack)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(it):
    proj_tail_fwd(h1, h2, w, pack)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / it
print(f"N={n} K={k} D={d}: {ms * 1e3:.1f} us per call = {4.0 * n * k * d / ms / 1e9:.0f} TFLOP/s")
