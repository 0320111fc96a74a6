timeout 300 python -m pytest tests/test_gpu_graph.py -x -q 2>&1 | tail -15
