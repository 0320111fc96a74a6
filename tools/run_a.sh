set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python tools/pcie_probe.py 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err; echo bench rc=$?
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_A.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
        print([(p["D"], round(p["ms"], 4), round(p["roofline"]["frac_whole_call"], 3)) for p in d["loss_sweep"]["points"]])
PY
timeout 120 python tools/loss_only.py 128 8192 5 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:bt_stat_norm_small -c 1 -o gpurun_out/r2_stat_small2 -f python tools/loss_only.py 128 8192 5 2>&1 | tail -2
ncu -i gpurun_out/r2_stat_small2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=dict(zip(h,r))
    for k in ('gpu__time_duration.sum','dram__bytes_read.sum','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active'): print(k, d.get(k))
"
