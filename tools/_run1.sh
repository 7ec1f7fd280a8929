set -x
timeout 600 python -m pytest tests/test_gpu_darcy.py -x -q -m gpu 2>&1 | tail -5
CES_BENCH_TAG=_tile timeout 300 python tools/bench_darcy.py 2>&1 | tail -4
CES_DARCY_CLUSTER=8 CES_BENCH_TAG=_tile_c8 timeout 300 python tools/bench_darcy.py 2>&1 | tail -4
CES_DARCY_CLUSTER=2 CES_BENCH_TAG=_tile_c2 timeout 300 python tools/bench_darcy.py 2>&1 | tail -4
