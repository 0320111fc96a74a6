"""Per-stage device time of the multi-GPU objective (run under torchrun on a GPU box):
    python -m torch.distributed.run --nproc-per-node R tools/dist_profile.py [N_local] [D]"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssl_audio_b200 import dist as D

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Dm = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(rank)
z1 = torch.randn(N, Dm, device=dev, generator=g)
z2 = (0.6 * z1 + 0.8 * torch.randn(N, Dm, device=dev, generator=g)).bfloat16()
z1 = z1.bfloat16()
be = D._CUDA_BACKEND
begin, count = D.row_block(Dm, world, rank)
w = be.workspace(dev, N, world, Dm, count)
names = ["stats_local", "ag_pack", "normalize", "ag_zh", "rows", "a2a", "allreduce"]
acc = {k: 0.0 for k in names}
iters = 20
for it in range(iters + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    be.stats_local(w, z1, z2, world, count); ev[1].record()
    dist.all_gather_into_tensor(w["pack_all"], w["pack_local"]); ev[2].record()
    be.normalize(w, z1, z2, world, rank, count, 1e-5, 0.1, None, None); ev[3].record()
    dist.all_gather_into_tensor(w["zh1"], w["zh1"][rank * N:(rank + 1) * N])
    dist.all_gather_into_tensor(w["zh2"], w["zh2"][rank * N:(rank + 1) * N]); ev[4].record()
    parts, a, b = be.rows(w, z1.dtype, dev, N, world, Dm, begin, count, 1.0, 0.005, False, float(world), 3); ev[5].record()
    outs = []
    for dzr in (a, b):
        recv = torch.empty((world, N, count), dtype=dzr.dtype, device=dev)
        dist.all_to_all_single(recv, dzr.view(world, N, count))
        outs.append(recv.permute(1, 0, 2).reshape(N, Dm))
    ev[6].record()
    off = parts[:2].clone(); dist.all_reduce(off); ev[7].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, n in enumerate(names):
            acc[n] += ev[k].elapsed_time(ev[k + 1])
if rank == 0:
    tot = sum(acc.values()) / iters
    print(f"world {world} N_local {N} D {Dm}: total {tot*1e3:.0f} us  " + "  ".join(f"{k} {v/iters*1e3:.0f}" for k, v in acc.items()))
dist.destroy_process_group()
