export CES_BENCH_SHAPES="128,256,2048,1.0"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:darcy_pcg_tile -c 1 -f -o gpurun_out/darcy_tile_r01c python tools/bench_darcy.py > gpurun_out/ncu_darcy.log 2>&1
tail -2 gpurun_out/ncu_darcy.log | cut -c1-200
