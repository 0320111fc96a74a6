"""What bounds `e2e`: PCIe.  Times, on one GPU, (a) a plain pinned-host -> device copy of the step's byte count, (b) the frontend fed from
pinned host waveforms (span gather over PCIe + log-mel + views), (c) the embedding upload, (d) b and c together as the e2e loop issues them."""
import os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
import bench as B

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = B._args_ns(8192) if hasattr(B, "_args_ns") else None
n, L, D = 1024, 160000, 8192
wav_h = (torch.randn(n, L) * 0.1).pin_memory()
z_h = [torch.randn(n, D).bfloat16().pin_memory() for _ in range(2)]
z_d = [torch.empty(n, D, dtype=torch.bfloat16, device=dev) for _ in range(2)]
fe = S.BatchFrontend(cfg, norm_stats=B.AS_STATS, path="lms")
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


blob_h = torch.empty(100 << 20, dtype=torch.uint8).pin_memory()
blob_d = torch.empty(100 << 20, dtype=torch.uint8, device=dev)
t = timed(lambda: blob_d.copy_(blob_h, non_blocking=True))
print(f"DMA H2D 100 MiB: {t:.3f} ms = {104.86 / t:.1f} GB/s")
t = timed(lambda: fe(wav_h))
hb = fe.h2d_bytes
print(f"frontend from pinned host ({hb / 1e6:.1f} MB over PCIe): {t:.3f} ms = {hb / t / 1e6:.1f} GB/s")
t = timed(lambda: [z_d[k].copy_(z_h[k], non_blocking=True) for k in range(2)])
print(f"embedding upload (33.6 MB): {t:.3f} ms = {33.55 / t:.1f} GB/s")


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        fe(wav_h)
    with torch.cuda.stream(s2):
        for k in range(2):
            z_d[k].copy_(z_h[k], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


t = timed(both)
print(f"both concurrently ({(hb + 33.55e6) / 1e6:.1f} MB): {t:.3f} ms = {(hb + 33.55e6) / t / 1e6:.1f} GB/s")
