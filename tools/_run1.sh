timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
CES_BENCH_TAG=_final4 timeout 300 python tools/bench_darcy.py 2>&1 | tail -4 | cut -c1-100
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
