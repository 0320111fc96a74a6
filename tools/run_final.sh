timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo ref rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_final.json", "gpurun_out/r2_bench_ref.json"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, "value", d["value"], "ms", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"), "launches", d.get("gpu_launches"))
            if "loss_sweep" in d:
                print([(p["D"], round(p["ms"], 4), round(p["roofline"]["frac_whole_call"], 3), round(p["roofline"]["frac"], 3)) for p in d["loss_sweep"]["points"]])
                print("sustained", d["sustained"]["value"], d["sustained"]["clocks"]["sm_mhz"], d["sustained"]["clocks"]["reasons"], "clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "proj", d["proj_tail"]["fused_ms"], d["proj_tail"]["chain_fwd_bwd_ms"])
PY
