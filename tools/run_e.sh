timeout 200 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_offline.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/frontend_only.py 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:logmel_kernel -c 6 python tools/frontend_only.py 2>&1 | grep -i "duration" | tail -4
