#!/bin/bash
# 8-GPU evidence run (one box): parity of the global-batch objective at 8 and 4 ranks through both exchange paths, a short tuning matrix
# (copy-engine exchange on/off x SM reservation), then the full bench line at 8 ranks.  Outputs: gpurun_out/r2_*; every command is bounded.
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
chk() {  # world tag extra-env...
  local R=$1 TAG=$2; shift 2
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $((29600 + R)) \
      tools/dist_check.py $BIG > gpurun_out/r2_dist_check_${R}_$TAG.log 2>&1
  echo "dist_check world=$R $TAG rc=$?"; grep "DIST_CHECK\|FAIL\|exchange" gpurun_out/r2_dist_check_${R}_$TAG.log | tail -3
}
BIG=--big chk 8 ce ABT_DIST_CE=1
BIG=      chk 8 nccl ABT_DIST_CE=0
BIG=--big chk 4 ce ABT_DIST_CE=1
run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 60 --warmup 10 --quick 2>/dev/null | grep quick; }
{
run ABT_DIST_CE=0 ABT_DIST_RESERVE_SMS=0
run ABT_DIST_CE=1 ABT_DIST_RESERVE_SMS=0
run ABT_DIST_CE=1 ABT_DIST_RESERVE_SMS=12
run ABT_DIST_CE=0 ABT_DIST_RESERVE_SMS=12
} > gpurun_out/r2_tune_n8.log 2>&1
cat gpurun_out/r2_tune_n8.log | cut -c1-330
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 \
    bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench world=8 rc=$?"; tail -2 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n8.json"))
    print("N=8 value", d["value"], "ms/step", d["ms_per_step"], "loss_fwd_bwd_ms", d["roofline"]["loss_fwd_bwd_ms"], "frac", d["roofline"]["frac"], "spread", d["loss_spread_over_ranks"])
    print("   sweep", [(p["D"], round(p["ms"], 4), round(p["roofline"]["frac"], 3)) for p in d["loss_sweep"]["points"]], "sustained", d["sustained"]["value"])
except Exception as e:
    print("bench N=8: no JSON line:", e)
PY
