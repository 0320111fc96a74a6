timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_D.json 2> gpurun_out/r2_bench_D.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_D.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_D.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
        print(json.dumps(d.get("proj_tail"), indent=1))
PY
