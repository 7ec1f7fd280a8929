"""Update-kernel sweep (BASELINE.json configs[4]): one EKS/ALDI step over (J, d, k) on N GPUs, reported as
particle-updates/s, algorithmic TFLOP/s (W_step of SURVEY.md 8d) and fraction of the nominal FP64 tensor peak.

    python tools/sweep.py [--quick]                      # 1 GPU
    torchrun --nproc-per-node 8 tools/sweep.py           # ensemble sharded by particle columns
Writes gpurun_out/sweep_n<N>.json.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200.engine import Engine, shard_range  # noqa: E402

NOMINAL = 148 * 128 * 1.965e9 / 1e12


def flops(J, d, k):
    return 2.0 * k * J * J + 2.0 * d * J * J + k * J + 6.0 * d * d * J + d ** 3 / 3.0


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    shapes = [(1024, 1024, 4096), (4096, 1024, 4096), (16384, 1024, 4096), (65536, 1024, 4096),
              (16384, 64, 256), (65536, 64, 256), (262144, 64, 256), (16384, 4096, 16384), (65536, 4096, 16384)]
    if "--quick" not in sys.argv:
        shapes.append((262144, 1024, 4096))
    if os.environ.get("CES_SWEEP_SHAPES"):           # "J,d,k;J,d,k"
        shapes = [tuple(int(v) for v in item.split(",")) for item in os.environ["CES_SWEEP_SHAPES"].split(";")]
    out = []
    for (J, d, k) in shapes:
        lo, hi = shard_range(J, rank, world)
        need = 8.0 * (hi - lo) * (3 * k + 8 * d) + 8.0 * J * (k + d) * (world > 1) + (8 << 30)
        if need > 150e9:
            continue
        gen = torch.Generator(device=dev).manual_seed(10 + rank)
        g0 = torch.Generator(device=dev).manual_seed(0)
        A = torch.randn(k, d, dtype=torch.float64, device=dev, generator=g0) / d ** 0.5
        y = torch.randn(k, dtype=torch.float64, device=dev, generator=g0)
        U = 10.0 * torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen)
        G = A @ U
        xi = torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen)
        res = torch.empty_like(U)
        eng = Engine(d, k, J, group=group)
        eng.set_problem(y.cpu().numpy(), 0.01 * np.eye(k), 100.0 * np.eye(d), np.zeros(d), np.zeros(d))
        steps = 1 if flops(J, d, k) / world > 2e14 else 3
        eng.step("aldi", U, G, xi, out=res)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.step("aldi", U, G, xi, out=res)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        tf = flops(J, d, k) / (ms * 1e-3) * 1e-12
        row = dict(J=J, d=d, k=k, n_gpus=world, ms_per_step=ms, particle_updates_per_s=J / (ms * 1e-3), tflops=tf,
                   frac_of_nominal_fp64_peak=tf / (NOMINAL * world))
        if rank == 0:
            print(row, flush=True)
        out.append(row)
        eng.close()
        del A, U, G, xi, res
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/sweep_n%d.json" % world, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
