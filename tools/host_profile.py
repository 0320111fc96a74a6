"""Where does the host time of one hot-path step go?  (run on a GPU box)"""
import os, sys, time, types, random
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
from bench import _args_ns, AS_STATS

B, D, L = 1024, 8192, 160000
dev = torch.device("cuda", 0)
cfg = _args_ns(D)
wav = 0.1 * torch.randn(B, L, device=dev)
z1 = torch.randn(B, D, device=dev).bfloat16(); z2 = torch.randn(B, D, device=dev).bfloat16()
fe = S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="crop")
crit = S.BarlowTwinsLoss(cfg, ncrops=2).to(dev)

def T(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    print(f"{name:40s} host {t_host*1e6:9.1f} us   host+gpu {t_all*1e6:9.1f} us")

eng = fe.transform.engine(B); eng.ensure_ring(dev)
T("planner.plan (host only)", lambda: eng.planner.plan(B, time_crop_range=905))
T("planner.plan + upload", lambda: eng.planner.plan(B, time_crop_range=905, device=dev))
T("frontend fe(wav)", lambda: fe(wav))
def loss_only():
    a = z1.detach().requires_grad_(True); b = z2.detach().requires_grad_(True)
    l = crit(b, a); l.backward()
T("loss fwd+bwd (module + autograd)", loss_only)
T("bt_loss_fwd_bwd (functional)", lambda: S.bt_loss_fwd_bwd(z1, z2, 1.0, 0.005, False))
def step():
    fe(wav); loss_only()
T("full step", step)

if os.environ.get("HOST_CPROFILE", "1") == "1":
    import cProfile, pstats, io
    for name, fn in (("loss fwd+bwd", loss_only), ("frontend", lambda: fe(wav))):
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            fn()
        pr.disable()
        torch.cuda.synchronize()
        out = io.StringIO()
        pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(22)
        print(f"==== cProfile of 200 x {name} (cumulative seconds / 200 = per call)")
        print("\n".join(l[:170] for l in out.getvalue().splitlines()[4:40]))
