#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 100 --warmup 10 --quick 2>/dev/null | grep quick | cut -c1-260; }
run BENCH_PLAN_AHEAD=1
run BENCH_PLAN_AHEAD=0
run BENCH_PLAN_AHEAD=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n2_final.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r2_bench_n2_final.json") if l.startswith("{")][-1]
print("N=2 value", d["value"], "ms", d["ms_per_step"], "host", d["host_enqueue_ms_per_step"], "sustained", d["sustained"]["value"], d["sustained"]["ms_per_step"], "e2e", d["e2e"]["value"], "spread", d["loss_spread_over_ranks"])
PY
