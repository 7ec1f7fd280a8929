export CES_BENCH_SHAPES="128,256,2048,1.0"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:darcy_pcg_tile -c 1 -f -o gpurun_out/darcy_tile_r01b python tools/bench_darcy.py > gpurun_out/ncu_darcy.log 2>&1
tail -2 gpurun_out/ncu_darcy.log
unset CES_BENCH_SHAPES
timeout 600 python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_cfg2_plain.json 2>gpurun_out/b_cfg2_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log | cut -c1-300
CES_BENCH_TAG=_final timeout 300 python tools/bench_darcy.py 2>&1 | tail -4
