"""Parity of the multi-GPU objective against the single-process global-batch oracle (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node R tools/dist_check.py
Exercises the native one-call step (abt_bt_dist_step) and the torch.distributed choreography (ABT_DIST_C10D=1 path)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import abt_oracle as O
from ssl_audio_b200 import dist as D
from ssl_audio_b200 import _lib

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
for (n, d, hsic, dt, need) in [(64, 256, False, torch.float32, (True, True)), (48, 512, True, torch.float32, (True, True)), (96, 1024, True, torch.float32, (True, True)),
                                (128, 2048, False, torch.bfloat16, (True, True)), (64, 256, False, torch.float32, (False, True))]:
    if d % (8 * world):
        continue
    z1g, z2g = O.synth_embeddings(world * n, d, seed=n + d)
    z1 = torch.from_numpy(z1g[rank * n:(rank + 1) * n]).to(dev).to(dt)
    z2 = torch.from_numpy(z2g[rank * n:(rank + 1) * n]).to(dev).to(dt)
    rl, r1, r2, _ = O.bt_loss_forward_backward(z1g, z2g, 1.0, 0.005, hsic)
    for mode in ("native", "plain", "c10d"):
        os.environ["ABT_DIST_C10D"] = "1" if mode == "c10d" else "0"
        _lib.load().abt_debug_set(7, 0 if mode == "plain" else 1)       # plain = native step without the exchange schedule
        rm, rv = torch.zeros(d, device=dev), torch.ones(d, device=dev)
        hook_calls = []
        loss, dz1, dz2 = D.bt_loss_fwd_bwd_global(z1, z2, 1.0, 0.005, hsic, running_mean=rm, running_var=rv, need_dz1=need[0], need_dz2=need[1],
                                                  grad_scale=1.0, overlap_hook=lambda: hook_calls.append(1))
        torch.cuda.synchronize()
        slack = 0.0 if dt == torch.float32 else 4e-3
        e_loss = abs(float(loss) - rl) / abs(rl)
        e1 = rel(dz1.float().cpu().numpy(), r1[rank * n:(rank + 1) * n]) if need[0] else 0.0
        e2 = rel(dz2.float().cpu().numpy(), r2[rank * n:(rank + 1) * n]) if need[1] else 0.0
        m, v = O.bn_running_update(np.zeros(d), np.ones(d), z1g)
        m, v = O.bn_running_update(m, v, z2g)
        e_rm = float(np.abs(rm.cpu().numpy() - m).max())
        good = e_loss < 1e-3 and e1 < 1e-3 + slack and e2 < 1e-3 + slack and e_rm < 1e-4 and len(hook_calls) == 1 and (dz1 is None) == (not need[0])
        ok = ok and good
        print(f"[rank {rank}] {mode:6s} n={n} d={d} hsic={hsic} {str(dt)[6:]} need={need}: loss {e_loss:.1e} dz1 {e1:.1e} dz2 {e2:.1e} rm {e_rm:.1e} {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
