timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_final.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"], "host", d["host_enqueue_ms_per_step"], d["step_launch"], "fe", d["frontend"]["ms_per_step"], d["frontend"]["hbm_frac"], "corr/grad", d["roofline"]["corr_ms"], d["roofline"]["grad_ms"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
PY
