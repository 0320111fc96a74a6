#!/bin/bash
# 8 ranks: does gating the frontend kernels on the objective's stream position (BENCH_HOOK_WAIT) hide them behind the embedding gathers?
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 60 --warmup 10 --quick 2>/dev/null | grep quick; }
{
run BENCH_HOOK_WAIT=1
run BENCH_HOOK_WAIT=0
run BENCH_HOOK_WAIT=1 ABT_DIST_RESERVE_SMS=0
run BENCH_HOOK_WAIT=1
} > gpurun_out/r2_tune_n8b.log 2>&1
cut -c1-200 gpurun_out/r2_tune_n8b.log
