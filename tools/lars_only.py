"""LARS / EMA alone on a model-sized parameter set: python tools/lars_only.py  -> ms per step and GB/s (5 fp32 passes per element)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
dev = torch.device("cuda", 0)
shapes = [(8192, 2048), (8192,), (8192, 8192), (8192,), (8192, 8192), (2048, 512), (2048,), (2048, 2048), (64, 64, 3, 3)]
params = [torch.nn.Parameter(torch.randn(*s, device=dev) * 0.02) for s in shapes]
for p in params:
    p.grad = 1e-3 * torch.randn_like(p)
opt = S.LARS(params, lr=0.2, weight_decay=1e-6, weight_decay_filter=True, lars_adaptation_filter=True)
n = sum(p.numel() for p in params)
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    opt.step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"LARS step over {n / 1e6:.1f} M parameters: {ms:.3f} ms = {7 * 4 * n / ms / 1e6:.0f} GB/s (2 + 3 reads, 2 writes per element)")
