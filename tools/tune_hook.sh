#!/bin/bash
# R ranks: frontend placement variants of the multi-GPU step (gated hook / ungated hook / no hook), quick bench each
R=$1
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $R --steps 100 --warmup 10 --quick 2>/dev/null | grep quick; }
{
run BENCH_HOOK_WAIT=1
run BENCH_HOOK_WAIT=0
run BENCH_HOOK=0
run BENCH_HOOK_WAIT=1
} > gpurun_out/r2_tune_hook_n$R.log 2>&1
cut -c1-220 gpurun_out/r2_tune_hook_n$R.log
