"""Pure host time of each half of a hot-path step: every call is timed alone, with the device idle before it and no synchronisation
inside the timed window, so neither the launch queue nor the plan staging ring can pace the host.  (run on a GPU box)
usage: host_isolated.py [calls] [--profile]"""
import cProfile, os, pstats, sys, time, random
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
from bench import _args_ns, AS_STATS

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 40
B, D, L = 1024, 8192, 160000
dev = torch.device("cuda", 0)
cfg = _args_ns(D)
wav = 0.1 * torch.randn(B, L, device=dev)
z1 = torch.randn(B, D, device=dev).bfloat16(); z2 = torch.randn(B, D, device=dev).bfloat16()
np.random.seed(0); random.seed(0)
fe = S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="crop")
crit = S.BarlowTwinsLoss(cfg, ncrops=2).to(dev)


def loss_step():
    a = z1.detach().requires_grad_(True); b = z2.detach().requires_grad_(True)
    crit(b, a, ngcrops_each=1).backward()


def isolated(name, fn):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    ts = np.array(ts) * 1e6
    print(f"{name:34s} host median {np.median(ts):7.1f} us   p10 {np.percentile(ts, 10):7.1f}   p90 {np.percentile(ts, 90):7.1f}")


isolated("frontend fe(wav)", lambda: fe(wav))
isolated("frontend prepare (plan + upload)", lambda: fe.prepare(wav))
h = fe.prepare(wav)
isolated("frontend launch (2 kernels)", lambda: fe.launch(h))
isolated("loss module fwd + backward", loss_step)
isolated("bt_loss_fwd_bwd (functional)", lambda: S.bt_loss_fwd_bwd(z1, z2, 1.0, 0.005, False))
isolated("full step", lambda: (fe(wav), loss_step()))
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    for _ in range(n):
        torch.cuda.synchronize()
        pr.enable(); fe(wav); loss_step(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
