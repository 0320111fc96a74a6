"""Parity of the multi-GPU objective against the single-process global-batch oracle (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node R tools/dist_check.py [--big]
Exercises the native one-call step (abt_bt_dist_step) with the exchange schedule forced on ("xchg") and off ("plain"), and the
torch.distributed choreography (ABT_DIST_C10D=1, "c10d"), plus the module-level multi-crop pairing (several calls per step).
fp32 outputs are held to 1e-3 against the oracle; bf16 outputs must be the rounding of the fp32 ones (2e-4).
--big adds BASELINE config 3 at full size (128 rows per rank, D = 4096 HSIC and D = 8192).  Rank 0 evaluates the float64 oracle
and broadcasts it.  The last line is `DIST_CHECK PASS` or `DIST_CHECK FAIL`."""
import os, sys, types
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import abt_oracle as O
from ssl_audio_b200 import dist as D
from ssl_audio_b200 import _lib
import ssl_audio_b200 as S

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
TOL, ROUND_TOL = 1e-3, 2e-4


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def oracle_on_rank0(fn, shapes):
    """Evaluate the float64 oracle on rank 0 only and broadcast the arrays (8 ranks share the host's cores)."""
    outs = fn() if rank == 0 else None
    res = []
    for k, shp in enumerate(shapes):
        t = torch.from_numpy(np.ascontiguousarray(outs[k], dtype=np.float64)).to(dev) if rank == 0 else torch.empty(shp, dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        res.append(t.cpu().numpy())
    return res


def set_mode(mode):
    os.environ["ABT_DIST_C10D"] = "1" if mode == "c10d" else "0"
    _lib.load().abt_debug_set(7, 0 if mode == "plain" else 1)


cases = [(64, 256, False, torch.float32, (True, True)), (48, 512, True, torch.float32, (True, True)), (96, 1024, True, torch.float32, (True, True)),
         (128, 2048, False, torch.bfloat16, (True, True)), (64, 256, False, torch.float32, (False, True)), (160, 1024, False, torch.bfloat16, (True, True))]
if "--big" in sys.argv:
    cases += [(128, 4096, True, torch.bfloat16, (True, True)), (128, 8192, False, torch.bfloat16, (True, True))]
for (n, d, hsic, dt, need) in cases:
    if d % (64 * world):
        continue
    z1g, z2g = O.synth_embeddings(world * n, d, seed=n + d)
    sl = slice(rank * n, (rank + 1) * n)
    rl, r1, r2 = oracle_on_rank0(lambda: (lambda t: (np.array([t[0]]), t[1], t[2]))(O.bt_loss_forward_backward_blocked(z1g, z2g, 1.0, 0.005, hsic)),
                                 [(1,), (world * n, d), (world * n, d)])
    rl = float(rl[0])
    m, v = O.bn_running_update(np.zeros(d), np.ones(d), z1g)
    m, v = O.bn_running_update(m, v, z2g)
    for mode in ("xchg", "plain", "c10d"):
        set_mode(mode)
        res = {}
        for tdt in ([torch.float32] if dt == torch.float32 else [torch.float32, dt]):
            z1 = torch.from_numpy(z1g[sl]).to(dev).to(tdt)
            z2 = torch.from_numpy(z2g[sl]).to(dev).to(tdt)
            rm, rv = torch.zeros(d, device=dev), torch.ones(d, device=dev)
            hook_calls = []
            loss, dz1, dz2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, running_mean=rm, running_var=rv, need_dz1=need[0], need_dz2=need[1],
                                                      grad_scale=1.0, overlap_hook=lambda: hook_calls.append(1))
            torch.cuda.synchronize()
            res[tdt] = (float(loss), dz1.float().cpu().numpy() if need[0] else None, dz2.float().cpu().numpy() if need[1] else None,
                        float(np.abs(rm.cpu().numpy() - m).max()), len(hook_calls))
        l32, a1, a2, e_rm, hooks = res[torch.float32]
        e_loss = abs(l32 - rl) / abs(rl)
        e1 = rel(a1, r1[sl]) if need[0] else 0.0
        e2 = rel(a2, r2[sl]) if need[1] else 0.0
        good = e_loss < TOL and e1 < TOL and e2 < TOL and e_rm < 1e-4 and hooks == 1 and (a1 is None) == (not need[0])
        extra = ""
        if dt != torch.float32:
            l16, b1, b2, e_rm16, _ = res[dt]
            q1 = rel(b1, torch.from_numpy(a1).to(dt).float().numpy().astype(np.float64))
            q2 = rel(b2, torch.from_numpy(a2).to(dt).float().numpy().astype(np.float64))
            good = good and abs(l16 - rl) / abs(rl) < TOL and q1 < ROUND_TOL and q2 < ROUND_TOL and e_rm16 < 1e-4
            extra = f" | {str(dt)[6:]} out vs rounded fp32 out: {q1:.1e} {q2:.1e}"
        ok = ok and good
        print(f"[rank {rank}/{world}] {mode:5s} n={n} d={d} hsic={hsic} need={need}: loss {e_loss:.1e} dz1 {e1:.1e} dz2 {e2:.1e} rm {e_rm:.1e}{extra} "
              f"{'OK' if good else 'FAIL'}", flush=True)

# ---- several objective calls per step through the module (multi-crop pairing of utils/loss.py:32-48: ncrops = 3 -> two terms)
n, d = 64, 512
if d % (64 * world) == 0:
    rng = np.random.default_rng(77)
    teacher_g = O.round_bf16(rng.standard_normal((world * n, d)).astype(np.float32))
    stud_g = [O.round_bf16((0.6 * teacher_g + 0.8 * rng.standard_normal((world * n, d))).astype(np.float32)) for _ in range(2)]
    sl = slice(rank * n, (rank + 1) * n)
    rl, rds, rdt = O.bt_loss_multicrop(np.concatenate(stud_g, 0), teacher_g, ncrops=3, ngcrops_each=1)
    for mode in ("xchg", "plain"):
        set_mode(mode)
        cfg = types.SimpleNamespace(projector_out_dim=d, HSIC=False, alpha=1.0, lmbda=0.005)
        crit = S.BarlowTwinsLoss(cfg, ncrops=3).to(dev)
        crit.grad_scale = 1.0
        student = torch.from_numpy(np.concatenate([s[sl] for s in stud_g], 0)).to(dev).requires_grad_(True)
        teacher = torch.from_numpy(teacher_g[sl]).to(dev).requires_grad_(True)
        loss = crit(student, teacher, ngcrops_each=1)
        loss.backward()
        torch.cuda.synchronize()
        ref_ds = np.concatenate([rds[v * world * n:(v + 1) * world * n][sl] for v in range(2)], 0)
        e_loss = abs(float(loss) - rl) / abs(rl)
        e1, e2 = rel(student.grad.cpu().numpy(), ref_ds), rel(teacher.grad.cpu().numpy(), rdt[sl])
        good = e_loss < TOL and e1 < TOL and e2 < TOL
        ok = ok and good
        print(f"[rank {rank}/{world}] {mode:5s} module ncrops=3 n={n} d={d}: loss {e_loss:.1e} dstudent {e1:.1e} dteacher {e2:.1e} {'OK' if good else 'FAIL'}", flush=True)
_lib.load().abt_debug_set(7, -1)
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DIST_CHECK world={world}", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
