"""Timing of the batched Darcy forward model (BASELINE configs 2 and 4 shapes) on one B200."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200 import darcy as cdarcy  # noqa: E402
from ces_b200.engine import Engine  # noqa: E402

out = []
SHAPES = [(64, 64, 1024, 1.0), (64, 64, 1024, 10.0), (128, 256, 8192, 1.0), (128, 256, 8192, 10.0)]
if os.environ.get("CES_BENCH_SHAPES"):      # "N,p,J,scale;N,p,J,scale"
    SHAPES = [(int(a), int(b), int(c), float(d)) for a, b, c, d in
              (item.split(",") for item in os.environ["CES_BENCH_SHAPES"].split(";"))]
for (N, p, J, scale) in SHAPES:
    m = cdarcy.model_trunc(Nmesh=N, p=p)
    m.max_iter = int(os.environ.get('CES_BENCH_MAXITER', '0'))
    rng = np.random.default_rng(0)
    m.obs_index = rng.choice(N * N, size=50, replace=False)
    U = torch.from_numpy(scale * rng.standard_normal((p, J))).cuda()
    G = torch.empty(50, J, dtype=torch.float64, device="cuda")
    eng = Engine(p, 50, J)
    def run():
        try:
            m.evaluate_ensemble(eng, U, G)
        except Exception as exc:          # a capped max_iter (timing experiments) reports non-convergence
            if not m.max_iter:
                raise
            m.last_iterations = m.max_iter
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    members_, total_its_, solver_ms_ = m.last_stats()
    r = dict(mean_iterations=total_its_ / max(members_, 1), solver_ms=solver_ms_, cluster=os.environ.get('CES_DARCY_CLUSTER', 'auto'),
             N=N, p=p, J=J, prior_scale=scale, ms=ms, members_per_s=J / ms * 1e3, cg_iterations=m.last_iterations)
    print(r, flush=True)
    out.append(r)
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("CES_BENCH_TAG", "")
json.dump(out, open("gpurun_out/bench_darcy%s.json" % tag, "w"), indent=1)
