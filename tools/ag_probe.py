"""All-gather bandwidth probe (torchrun): NCCL all_gather_into_tensor vs torch symmetric-memory multimem all-gather."""
import os, sys, time
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, D = 1024, 8192
def timeit(fn, iters=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
out = torch.empty((world * N, D), dtype=torch.float16, device=dev)
src = out[rank * N:(rank + 1) * N]
src.normal_()
t = timeit(lambda: dist.all_gather_into_tensor(out, src))
if rank == 0: print(f"nccl all_gather in-place 16MB/rank: {t:.0f} us  ({(world-1)*N*D*2/t/1e3:.0f} GB/s ingress)  NCCL_ALGO={os.environ.get('NCCL_ALGO')}")
big = torch.empty((2, world * N, D), dtype=torch.float16, device=dev)
def two():
    dist.all_gather_into_tensor(big[0], big[0][rank * N:(rank + 1) * N]); dist.all_gather_into_tensor(big[1], big[1][rank * N:(rank + 1) * N])
t = timeit(two)
if rank == 0: print(f"nccl 2 x all_gather: {t:.0f} us  ({2*(world-1)*N*D*2/t/1e3:.0f} GB/s ingress)")
a = torch.empty((world * N, D // world), dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
t = timeit(lambda: dist.all_to_all_single(b, a))
if rank == 0: print(f"nccl all_to_all 16MB: {t:.0f} us")
try:
    import torch.distributed._symmetric_memory as symm
    sm = symm.empty((world * N, D), dtype=torch.float16, device=dev)
    hdl = symm.rendezvous(sm, dist.group.WORLD.group_name)
    sm[rank * N:(rank + 1) * N].normal_()
    if rank == 0: print("symm_mem ok; multicast:", getattr(hdl, "multicast_ptr", None) not in (None, 0))
    def mm():
        torch.ops.symm_mem.multimem_all_gather_out(sm[rank * N:(rank + 1) * N], dist.group.WORLD.group_name, sm)
    try:
        t = timeit(mm)
        if rank == 0: print(f"symm_mem multimem_all_gather_out: {t:.0f} us  ({(world-1)*N*D*2/t/1e3:.0f} GB/s ingress)")
    except Exception as e:
        if rank == 0: print("multimem_all_gather_out failed:", repr(e)[:300])
    # peer pull with copy engines: each rank copies the other ranks' slots out of their symmetric buffers
    peers = [hdl.get_buffer(r, (world * N, D), torch.float16) for r in range(world)]
    def pull():
        hdl.barrier(channel=0)
        for k in range(1, world):
            q = (rank + k) % world
            out[q * N:(q + 1) * N].copy_(peers[q][q * N:(q + 1) * N], non_blocking=True)
        hdl.barrier(channel=1)
    t = timeit(pull)
    if rank == 0: print(f"symm_mem peer pull (copy_ per peer): {t:.0f} us  ({(world-1)*N*D*2/t/1e3:.0f} GB/s ingress)")
except Exception as e:
    if rank == 0: print("symm_mem unavailable:", repr(e)[:300])
dist.destroy_process_group()
