"""Per-phase, per-rank CUDA-event timeline of one update (profiles/r02_timeline_*.json).

    python tools/timeline.py --workload target|cfg3 [--host]           (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/timeline.py --workload target

Every rank enables the library timeline (ces_timeline_*), runs warm-up steps and then `--reps` recorded steps, and
writes the marks (name, ms since the step's first mark) of the median step.  --host times the reference-facing host
call (ces_step_host / the sharded host step) instead of the device-resident step."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = {"target": (1024, 4096, 65536), "cfg3": (1024, 4096, 16384), "small": (64, 50, 1024)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="target", choices=sorted(SHAPES))
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ces_b200.engine import Engine, shard_range

    world, rank, local = (int(os.environ.get(v, d)) for v, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    d, k, J = SHAPES[args.workload]
    lo, hi = shard_range(J, rank, world)
    gen = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=gen)
    A = rn(k, d) / d ** 0.5
    ustar = rn(d)
    y = A @ ustar + 0.1 * rn(k)
    gen_c = torch.Generator(device=dev).manual_seed(100 + rank)
    U = 10.0 * torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen_c)
    G = A @ U
    xi = torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen_c)
    out = torch.empty_like(U)
    eng = Engine(d, k, J, group=group)
    eng.set_problem(y.cpu().numpy(), 0.01 * np.eye(k), 100.0 * np.eye(d), np.zeros((d, 1)), ustar.cpu().numpy())
    if args.host:
        pin = lambda t: torch.empty(t.shape, dtype=torch.float64).pin_memory().copy_(t).numpy()
        Un, Gn, xn = pin(U), pin(G), pin(xi)
        step = lambda: eng.step_host("aldi", Un, Gn, xn)
    else:
        step = lambda: eng.step("aldi", U, G, xi, out=out)
    for _ in range(3):
        step()
    runs = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.timeline(True)
        eng.mark("step:begin")
        step()
        eng.mark("step:end")
        runs.append(eng.timeline_read())
        eng.timeline(False)
    runs.sort(key=lambda r: max(t for _, t in r))
    marks = runs[len(runs) // 2]
    name = "r02_timeline_%s_n%d%s_rank%d.json" % (args.workload, world, "_host" if args.host else "", rank)
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, name), "w") as fh:
        json.dump({"workload": args.workload, "d": d, "k": k, "J": J, "n_gpus": world, "rank": rank, "host_call": args.host,
                   "marks_ms": marks, "all_runs_total_ms": [max(t for _, t in r) for r in runs]}, fh, indent=1)
    if rank == 0:
        prev = 0.0
        for nm, t in sorted(marks, key=lambda m: m[1]):
            print("%-28s %10.3f ms  (+%.3f)" % (nm, t, t - prev))
            prev = t
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
