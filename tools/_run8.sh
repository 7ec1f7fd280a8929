timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n8.json 2> gpurun_out/bench_cfg4_n8.err
tail -c 600 gpurun_out/bench_cfg4_n8.json; tail -3 gpurun_out/bench_cfg4_n8.err
