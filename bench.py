#!/usr/bin/env python
"""bench.py -- particle-updates/s of one EKS (ALDI) step on B200, the headline metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|cfg3|small] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one ensemble Kalman update (sampling.eks_update_aldi,
ces/calibrate.py:451-490) of the synthetic linear-Gaussian ensemble of SURVEY.md section 8(d).

  value      J * steps / s with U, G, xi already resident in HBM (Engine.step, device pointers);
             timed with CUDA events on the stream the library launches on, max over ranks.
  e2e        the same metric through the reference-facing call sampling.eks_update_aldi(numpy arrays):
             pinned host buffers, host->device copies of U, G, xi and the device->host copy of U_next are
             inside the timed region (ces_step_host).  N > 1: each rank runs its column shard through the
             same phases with its own host<->device copies.
  roofline   the dominant kernel (D = (1/J) E^T W, FP64 DMMA): algorithmic flops / CUDA-event duration of
             its launches inside the timed region (ces_profile_*), against the FP64 tensor peak.  The
             MEASURED_PEAKS.json file holds no FP64 figure, so the denominator is the cuBLAS DGEMM rate
             measured in this run (torch.matmul fp64, yardstick only); nominal 148 SM x 128 flop/clk x
             1.965 GHz = 37.2 TF/s is reported beside it.
  cpu_baseline  the numpy restatement of the reference step as written (oracle/, kind "port": the reference
             itself is not importable on the GPU box) on the host cores, on a bounded sample.
  --impl reference  times only that CPU arm, K steps of the bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (d, k, J, J of the bounded CPU sample)
    "target": (1024, 4096, 65536, 2048),     # BASELINE.json "Target": one EKS step at J=65536, d=1024, k=4096
    "cfg3": (1024, 4096, 16384, 2048),       # BASELINE.json configs[2]
    "small": (64, 50, 1024, 1024),           # configs[1] shape (smoke-sized)
    "cfg1": (2, 10, 100, 100),               # configs[0]: the reference's own CPU-runnable case
}
# One EKS iteration INCLUDING the batched Darcy forward solve (ces/darcy.py model_trunc): (grid, d, n_obs, J, update rule)
DARCY_WORKLOADS = {
    "cfg2": (64, 64, 50, 1024, "eki"),       # BASELINE.json configs[1]: EKI, 64 x 64 grid, d = 64 KL modes, J = 1024
    "cfg4": (128, 256, 50, 65536, "aldi"),   # configs[3]: EKS, 128 x 128 grid, d = 256, J = 65536 (8192 per GPU at N = 8)
}
METRIC = "particle-updates/sec (J*steps/s) for one EKS step"
NOMINAL_FP64_TFLOPS = 148 * 128 * 1.965e9 / 1e12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("CES_BENCH_WORKLOAD", "target"), choices=sorted(list(WORKLOADS) + list(DARCY_WORKLOADS)))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--formulation", default="interaction", choices=["interaction", "factored"],
                    help="'interaction' forms the J x J matrix D like the reference (the graded formulation, default); "
                         "'factored' is the opt-in algorithmically reduced path (same update to rounding, D never formed)")
    return ap.parse_args()


def config_of(args, d, k, J, nranks):
    return {"workload": "EKS/ALDI step, synthetic linear-Gaussian, d=%d k=%d J=%d, Gamma=0.1^2 I, Sigma0=100 I (%s)"
                        % (d, k, J, args.workload),
            "d": d, "k": k, "J": J, "update": "aldi", "time_step": "default 1/(||D||_F+1e-8)",
            "parallelism": "particle columns sharded over %d GPU(s)" % nranks,
            "formulation": ("interaction: D = (1/J) E^T W formed in panels (reference formulation)"
                            if args.formulation == "interaction" else
                            "factored: ALGORITHMICALLY REDUCED, (U~ E^T) W and Gram-matrix ||D||_F, D never formed"),
            "l2": "inputs per step (U, G, xi = %.2f GB) exceed the 126 MB L2; no explicit flush" % (8.0 * J * (2 * d + k) / 1e9)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_rate(d, k, Js, steps, warmup):
    """particle-updates/s of the reference step as written (oracle port) at ensemble size Js."""
    from oracle import eks_oracle as eo

    pr = eo.linear_gaussian_problem(d, k, Js)
    args = (pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
    for _ in range(warmup):
        eo.step("aldi", *args, as_written=True)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        eo.step("aldi", *args, as_written=True)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return Js / t, t


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_baseline(d, k, J, Js, steps=1, warmup=1):
    from oracle import eks_oracle as eo

    rate, t = cpu_step_rate(d, k, Js, steps, warmup)
    full = rate * eo.reference_flops(Js, d, k) / Js / (eo.reference_flops(J, d, k) / J)
    return {"value": rate, "unit": "particle-updates/s", "cores": host_threads(), "kind": "port",
            "sample": "numpy restatement of sampling.eks_update_aldi as written (3 Gamma solves, 3 JxJ products), "
                      "d=%d k=%d at J=%d instead of J=%d, %.2f s/step; per-particle cost grows ~J so this "
                      "over-states the CPU at full J (flop-model extrapolation: %.3g particle-updates/s)"
                      % (d, k, Js, J, t, full),
            "extrapolated_full_J": full}


def darcy_problem(N, d, n_obs):
    """The scenario of examples/scripts/darcy-flow.py:9-36 on model_trunc: truth drawn with set_initial(seed=1), n_obs
    observation cells, gamma = 0.005, prior N(0, 100 I).  Host-side constants only; no solve happens here."""
    rng = np.random.default_rng(0)
    obs_index = np.sort(rng.choice(N * N, size=n_obs, replace=False))
    return {"obs_index": obs_index, "Gamma": 0.005 ** 2 * np.eye(n_obs), "mu": np.zeros((d, 1)),
            "Sigma0": 100.0 * np.eye(d)}


def darcy_cpu_rate(N, d, n_obs, J, rule, members, steps=1):
    """particle-updates/s of one host core running the scipy restatement of the Darcy solve (oracle/darcy_oracle.py, one
    sparse direct solve per member like the reference's MATLAB call) plus the numpy update at a bounded ensemble."""
    from oracle import darcy_oracle as do
    from oracle import eks_oracle as eo

    pr = darcy_problem(N, d, n_obs)
    model = do.ModelTrunc(Nmesh=N, p=d)
    model.obs_index = pr["obs_index"]
    model.set_initial(seed=1)
    rng = np.random.default_rng(1)
    U = 10.0 * rng.standard_normal((d, members))
    y = model(model.ustar) + 0.005 * rng.standard_normal(n_obs)
    ts = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        G = np.stack([model(U[:, j]) for j in range(members)], axis=1)
        xi = rng.standard_normal((d, members))
        eo.step("aldi" if rule != "eki" else "eki", y, U, G, pr["Gamma"], pr["mu"], pr["Sigma0"],
                model.ustar.reshape(d, 1), xi, as_written=(rule != "eki"))
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return members / t, t


def darcy_arm(args, dev, world, rank, local, group):
    """--workload cfg2 | cfg4: one ensemble Kalman iteration = batched Darcy forward solve of every member + update."""
    import torch
    import torch.distributed as dist
    from ces_b200 import calibrate
    from ces_b200 import darcy as cdarcy
    from ces_b200.engine import Engine, shard_range

    N, d, n_obs, J, rule = DARCY_WORKLOADS[args.workload]
    pr = darcy_problem(N, d, n_obs)
    lo, hi = shard_range(J, rank, world)
    model = cdarcy.model_trunc(Nmesh=N, p=d)
    model.obs_index = pr["obs_index"]
    model.set_initial(seed=1)
    model.n_obs = n_obs
    ustar = np.asarray(model.ustar, dtype=np.float64).reshape(d, 1)
    eng = Engine(d, n_obs, J, group=group)
    rng = np.random.default_rng(1)
    y = model(model.ustar) + 0.005 * rng.standard_normal(n_obs)         # truth through the device solver
    eng.set_problem(y, pr["Gamma"], pr["Sigma0"], pr["mu"], ustar)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    U = 10.0 * torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen)      # darcy-flow.py:87
    xi = torch.randn(d, hi - lo, dtype=torch.float64, device=dev, generator=gen)
    G = torch.empty(n_obs, hi - lo, dtype=torch.float64, device=dev)
    out = torch.empty_like(U)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    stats = {"iters": 0, "members": 0, "ms": 0.0}

    def step_dev():
        model.evaluate_ensemble(eng, U, G)
        m_, it_, ms_ = model.last_stats()
        stats["iters"] += it_
        stats["members"] += m_
        stats["ms"] += ms_
        eng.step(rule, U, G, None if rule == "eki" else xi, out=out)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    stats.update(iters=0, members=0, ms=0.0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = eng.launch_count()
    t0 = time.time()
    ms = timed(step_dev, args.steps)
    t1 = time.time()
    launches = eng.launch_count() - n0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = ms / args.steps

    # ---- end to end: sampling.run for ONE iteration from host arrays (H2D of U0, forward, update, the final forward
    # that run() always does, D2H of Ustar and Gstar); device Philox noise, no trace
    s = calibrate.sampling(d, n_obs, J)
    s.mu, s.sigma, s.ustar, s.T = pr["mu"], pr["Sigma0"], ustar, 1
    s.mute_bar = True
    if group is not None:
        s.group = group
    U0_h = 10.0 * np.random.default_rng(2).standard_normal((d, J))

    def step_e2e():
        if hasattr(s, "metrics"):
            del s.metrics
        s.run(y, U0_h, model, pr["Gamma"], None, trace=False, update=rule, rng="device", seed=1, t_tol=1e30)

    step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    if rank != 0:
        return
    nodes = (N - 2) * (N - 2)
    flops = 18.0 * nodes * stats["iters"]                       # 9 FMA per node and CG iteration
    achieved = flops / (stats["ms"] * 1e-3) * 1e-12 if stats["ms"] > 0 else None
    peak = 148 * 64 * 2 * 1.965e9 / 1e12
    roofline = {"bound": "fp64-vector (the solver keeps a member in registers/shared memory of its cluster: neither an "
                         "HBM nor a tensor-core kernel; contract enum does not fit)",
                "kernel": "darcy_pcg_tile_kernel (Jacobi-preconditioned CG, one cluster per member)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": "nominal FP64 FMA rate 148 SM x 64 FMA/clk x 1.965 GHz (no measured FP64 entry in MEASURED_PEAKS.json)",
                "share_of_step": stats["ms"] / ms if ms > 0 else None,
                "cg_iterations_mean": stats["iters"] / max(stats["members"], 1),
                "node_iterations_per_s": nodes * stats["iters"] / (stats["ms"] * 1e-3) if stats["ms"] > 0 else None,
                "algorithmic_hbm_bytes": 16.0 * N * N * stats["members"],
                "hbm_gbs": 16.0 * N * N * stats["members"] / (stats["ms"] * 1e-3) * 1e-9 if stats["ms"] > 0 else None,
                "traffic": None}
    line = {"metric": METRIC + " including the batched Darcy forward solve", "value": J / (ms_per_step * 1e-3),
            "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s iteration with batched Darcy forward (model_trunc Nmesh=%d, p=%d, n_obs=%d), J=%d (%s)"
                                   % (rule.upper(), N, d, n_obs, J, args.workload),
                       "d": d, "k": n_obs, "J": J, "grid": N, "update": rule,
                       "parallelism": "particle columns sharded over %d GPU(s)" % world,
                       "l2": "per-step fields (3 x %.2f GB) exceed the 126 MB L2; no explicit flush" % (8.0 * N * N * min(J // world, 4096) / 1e9)},
            "e2e": {"value": J / (ms_e2e * 1e-3), "unit": "particle-updates/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 8 * d * J, "d2h_bytes_per_step": 8 * (d + n_obs) * J,
                    "api": "ces_b200.calibrate.sampling.run(T=1, trace=False, rng='device'): forward + update + the final "
                           "forward run() always performs"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        members = 8 if N > 64 else 32
        rate, t = darcy_cpu_rate(N, d, n_obs, J, rule, members)
        line["cpu_baseline"] = {"value": rate, "unit": "particle-updates/s", "cores": 1, "kind": "port",
                                "sample": "scipy restatement of the Darcy solve (one sparse direct solve per member, like "
                                          "the reference's MATLAB call) + numpy update, %d members instead of %d, %.2f s"
                                          % (members, J, t)}
    emit(line)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in DARCY_WORKLOADS:
        N, d, n_obs, J, rule = DARCY_WORKLOADS[args.workload]
        members = 8 if N > 64 else 32
        rate, t = darcy_cpu_rate(N, d, n_obs, J, rule, members, steps=max(1, args.steps))
        sample = "scipy restatement of the Darcy solve + numpy update, %d members instead of %d, %.2f s/step" % (members, J, t)
        emit({"impl": "reference", "metric": METRIC + " including the batched Darcy forward solve", "value": rate,
              "unit": "particle-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic", "config": {"workload": "%s (%s)" % (args.workload, rule), "d": d, "k": n_obs, "J": J, "grid": N},
              "cpu_baseline": {"value": rate, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample},
              "e2e": {"value": rate, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return
    d, k, J, Js = WORKLOADS[args.workload]
    rate, t = cpu_step_rate(d, k, Js, max(1, args.steps), max(0, min(args.warmup, 1)))
    from oracle import eks_oracle as eo
    full = rate * (eo.reference_flops(Js, d, k) / Js) / (eo.reference_flops(J, d, k) / J)
    sample = ("numpy restatement of sampling.eks_update_aldi as written, d=%d k=%d at J=%d instead of J=%d "
              "(bounded sample; %.2f s/step; flop-model extrapolation to full J: %.3g particle-updates/s)"
              % (d, k, Js, J, t, full))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "particle-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, d, k, J, args.gpus),
            "cpu_baseline": {"value": rate, "unit": "particle-updates/s", "cores": host_threads(), "kind": "port",
                             "sample": sample, "extrapolated_full_J": full},
            "e2e": {"value": rate, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ GPU arm
_REAL_STDOUT = None


def emit(line):
    """Print the one JSON line on the real stdout (see main: fd 1 is pointed at stderr while the run is in progress so
    that library chatter such as NCCL's version banner cannot get in front of it)."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from ces_b200 import calibrate
    from ces_b200.engine import Engine, shard_range
    from oracle import eks_oracle as eo   # flop model only (cpu_baseline leg below times the port)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if args.workload in DARCY_WORKLOADS:
        darcy_arm(args, dev, world, rank, local, group)
        if world > 1:
            dist.destroy_process_group()
        return
    d, k, J, Js = WORKLOADS[args.workload]
    lo, hi = shard_range(J, rank, world)

    # synthetic problem, generated on the device; every rank draws the same global data
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
    Gamma = 0.01 * np.eye(k)
    Sigma0 = 100.0 * np.eye(d)
    mu = np.zeros((d, 1))
    y_h, ustar_h = y.cpu().numpy(), ustar.cpu().numpy().reshape(d, 1)

    eng = Engine(d, k, J, group=group)
    eng.set_problem(y_h, Gamma, Sigma0, mu, ustar_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA-event time, max over ranks (ms)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- FP64 yardstick (cuBLAS DGEMM through torch.matmul), rank 0 only, before the timed region
    dgemm_tflops = None
    if rank == 0:
        n = 8192
        a, b = rn(n, 4096), rn(4096, n)
        c = torch.empty(n, n, dtype=torch.float64, device=dev)
        best = 1e30
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        dgemm_tflops = 2.0 * n * n * 4096 / best * 1e-9
        del a, b, c

    # ---- device-resident steps
    def step_dev():
        eng.step("aldi", U, G, xi, out=out, formulation=args.formulation)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    eng.profile(True)
    eng.profile_read()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = eng.launch_count()
    t0 = time.time()
    ms = timed(step_dev, args.steps)
    t1 = time.time()
    launches = eng.launch_count() - n0
    gemm_ms, gemm_n, gemm_flops = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = ms / args.steps
    value = J / (ms_per_step * 1e-3)

    # ---- end to end: the reference-facing call on pinned host arrays
    pin = lambda t: torch.empty(t.shape, dtype=torch.float64).pin_memory().copy_(t)
    U_h, G_h, xi_h = pin(U), pin(G), pin(xi)
    if world == 1:
        s = calibrate.sampling(d, k, J)
        s.mu, s.sigma, s.ustar = mu, Sigma0, ustar_h
        s._engine, s._engine_key = eng, (d, k, J, id(None))
        Un, Gn, xn = U_h.numpy(), G_h.numpy(), xi_h.numpy()

        def step_e2e():
            s.eks_update_aldi(y_h, Un, Gn, Gamma, 0, xi=xn, formulation=args.formulation)
    else:
        out_h = torch.empty(U.shape, dtype=torch.float64).pin_memory()
        Ud, Gd, xd = torch.empty_like(U), torch.empty_like(G), torch.empty_like(xi)

        def step_e2e():
            Ud.copy_(U_h, non_blocking=True)
            Gd.copy_(G_h, non_blocking=True)
            xd.copy_(xi_h, non_blocking=True)
            eng.step("aldi", Ud, Gd, xd, out=out, formulation=args.formulation)
            out_h.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

    step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    e2e = {"value": J / (ms_e2e * 1e-3), "unit": "particle-updates/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": 8 * J * (2 * d + k), "d2h_bytes_per_step": 8 * J * d,
           "api": "ces_b200.calibrate.sampling.eks_update_aldi(numpy arrays) -> ces_step_host" if world == 1
                  else "Engine.step per rank with pinned host<->device copies of its shard"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_step = eo.algorithmic_flops(J, d, k)
    if args.formulation == "factored":
        # flops of the reduced path: P1 and V (2dkJ each), two symmetric Gram matrices (k^2 J each), covariance,
        # prior and noise products, Cholesky
        flops_step = 4.0 * d * k * J + 2.0 * k * k * J + 6.0 * d * d * J + d ** 3 / 3.0 + k * J
    peak = dgemm_tflops
    achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else None
    roofline = {"bound": "tensor", "kernel": "gemm_dmma_kernel<A_KM,B_KN> (D = (1/J) E^T W, FP64 DMMA.8x8x4 + TMA)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": "cuBLAS DGEMM 8192x8192x4096 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": NOMINAL_FP64_TFLOPS,
                "launches": gemm_n, "ms_per_launch": gemm_ms / max(gemm_n, 1),
                "share_of_step": gemm_ms / ms if ms > 0 else None,
                "traffic": TRAFFIC_BYTES.get(args.workload),
                "step_achieved_tflops": flops_step / (ms_per_step * 1e-3) * 1e-12,
                "step_frac_of_peak": flops_step / (ms_per_step * 1e-3) * 1e-12 / (peak * world) if peak else None}
    if args.formulation == "factored":
        roofline.update({"kernel": "whole step, factored formulation (no single dominant kernel; D GEMM not launched)",
                         "achieved": roofline["step_achieved_tflops"], "frac": roofline["step_frac_of_peak"],
                         "launches": None, "ms_per_launch": None, "share_of_step": None, "traffic": None,
                         "algorithmically_reduced": True})
    line = {"metric": METRIC, "value": value, "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, d, k, J, world),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(d, k, J, Js)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the D GEMM from the committed ncu capture
# (profiles/); None until a capture for that workload exists.
TRAFFIC_BYTES = {
    # profiles/r01_gemm_raster_sweep.csv (group 16, serpentine K): 9.53 GB read + 2.14 GB written for the 16384 x 16384 x 4096 launch
    # (algorithmic: E 0.54 + W 0.54 + D 2.15 GB; the excess is E/W panels streamed once per wave of 148 tiles: a wave's
    # unique footprint, ~100 MB, fills the L2, so there is no cross-wave reuse; 3 % of HBM bandwidth, duration unchanged)
    "cfg3": 11.67e9,
    # profiles/r01_ncu_full_gemm_d_target.csv: 56.98 GB read + 8.59 GB written per 65536 x 16384 x 4096 panel launch
    # (algorithmic: E 2.15 + W panel 0.54 + D panel 8.59 GB), same mechanism, 3.3 % of HBM bandwidth; captured before
    # the serpentine traversal (-23 % reads at cfg3)
    "target": 65.57e9,
}

if __name__ == "__main__":
    main()
