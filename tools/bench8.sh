#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 \
    bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench world=8 rc=$?"; tail -2 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_n8.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("N=8 value", d["value"], "ms/step", d["ms_per_step"], "loss_fwd_bwd_ms", d["roofline"]["loss_fwd_bwd_ms"], "frac", d["roofline"]["frac"], "spread", d["loss_spread_over_ranks"])
        print("   sweep", [(p["D"], round(p["ms"], 4), round(p["roofline"]["frac"], 3)) for p in d["loss_sweep"]["points"]], "sustained", d["sustained"]["value"], "e2e", d["e2e"]["value"])
PY
