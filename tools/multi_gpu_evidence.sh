#!/bin/bash
# Multi-GPU evidence run (one box, NCCL over NVLink): parity of the global-batch objective against the float64 oracle at every
# world size given, then the hot-path bench at the same sizes.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
#   usage: [TAG=_suffix] tools/multi_gpu_evidence.sh "2 4 8" "2 8"      (world sizes for dist_check, world sizes for bench)
CHECKS=${1:-"2"}
BENCHES=${2:-"2"}
mkdir -p gpurun_out
for R in $CHECKS; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $((29600 + R)) \
      tools/dist_check.py --big > gpurun_out/r2_dist_check_$R${TAG}.log 2>&1
  echo "dist_check world=$R rc=$?"; grep "DIST_CHECK\|FAIL\|exchange" gpurun_out/r2_dist_check_$R${TAG}.log | tail -4
done
for R in $BENCHES; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $((29700 + R)) \
      bench.py --gpus $R --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n$R${TAG}.json 2> gpurun_out/r2_bench_n$R${TAG}.err
  echo "bench world=$R rc=$?"; tail -2 gpurun_out/r2_bench_n$R${TAG}.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2_bench_n$R${TAG}.json") if l.startswith("{")][-1]
    print("N=$R value", d["value"], "ms/step", d["ms_per_step"], "loss_fwd_bwd_ms", d["roofline"]["loss_fwd_bwd_ms"], "frac", d["roofline"]["frac"], "spread", d["loss_spread_over_ranks"])
    print("   sweep", [(p["D"], round(p["ms"], 4), round(p["roofline"]["frac"], 3)) for p in d["loss_sweep"]["points"]], "sustained", d["sustained"]["value"])
except Exception as e:
    print("bench N=$R: no JSON line:", e)
PY
done
