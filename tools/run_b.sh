set -x
timeout 400 python bench.py > gpurun_out/r2_bench_B.json 2> gpurun_out/r2_bench_B.err; echo bench rc=$?
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_B.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "host", d["host_enqueue_ms_per_step"])
        print([(p["D"], round(p["ms"], 4), round(p["roofline"]["frac_whole_call"], 3)) for p in d["loss_sweep"]["points"]])
PY
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-s 0.05 --sweep-iters 3 > gpurun_out/b_short.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-s 0.05 --sweep-iters 3 > gpurun_out/ncu_list.log 2>&1
echo ncu rc=$?; wc -l gpurun_out/r2_launches_step.csv
