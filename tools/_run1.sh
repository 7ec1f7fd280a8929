timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 1500 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; tail -c 1500 gpurun_out/bench_cfg4_n1.json; tail -3 gpurun_out/bench_cfg4_n1.err
