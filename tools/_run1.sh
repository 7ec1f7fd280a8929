timeout 900 python -m pytest tests/test_gpu_darcy.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_darcy.py 2>&1 | tail -4 | cut -c1-260
