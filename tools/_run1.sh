timeout 1500 python -m pytest tests/test_gpu_darcy.py -x -q -m gpu 2>&1 | tail -4
