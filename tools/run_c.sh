timeout 400 python bench.py --e2e-probe --no-cpu-baseline 2>&1 | grep e2e_probe
