"""Timing of the batched Lorenz forward models (ces_b200/csrc/lorenz.cu) on one B200."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200 import utils as cu  # noqa: E402

out = []


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
# Lorenz 63 as in lorenz63.ipynb: window 10 time units at 100 samples per unit, two windows
m = cu.lorenz63(l_window=10, freq=100)
t = np.arange(0, 20.0 + 1e-9, 0.01)
for J in (1024, 65536):
    U = torch.from_numpy(np.array([[28.0], [8.0 / 3]]) + 0.5 * rng.standard_normal((2, J))).cuda()
    W0 = torch.from_numpy(np.array([[1.0], [2.0], [25.0]]) + rng.standard_normal((3, J))).cuda()
    G = torch.empty(9, J, dtype=torch.float64, device="cuda")
    Wend = torch.empty_like(W0)
    ms = timed(lambda: m.evaluate_ensemble_pde(None, U, W0, t, G, Wend))
    steps = (len(t) - 1) * m.substeps
    out.append(dict(model="lorenz63", J=J, rk4_steps=steps, ms=ms, particle_steps_per_s=J * steps / ms * 1e3))
    print(out[-1], flush=True)
# Lorenz 96 with the class defaults: 36 slow x 10 fast, spin-up 10 + one window of 10 time units at 10 samples per unit
m96 = cu.lorenz96()
t = np.arange(0, 20.0 + 1e-9, 0.1)
for J in (128, 1024):
    U = torch.from_numpy(np.array([[1.0], [10.0], [np.log(10.0)], [10.0]]) * (1 + 0.02 * rng.standard_normal((4, J)))).cuda()
    x = rng.random((36, J)) * 15 - 5
    W0 = torch.from_numpy(np.concatenate([x, 0.05 * np.repeat(x, 10, axis=0)], axis=0)).cuda()
    G = torch.empty(180, J, dtype=torch.float64, device="cuda")
    Wend = torch.empty_like(W0)
    ms = timed(lambda: m96.evaluate_ensemble_pde(None, U, W0, t, G, Wend), reps=2)
    steps = (len(t) - 1) * m96.substeps
    out.append(dict(model="lorenz96", J=J, rk4_steps=steps, ms=ms, particle_steps_per_s=J * steps / ms * 1e3,
                    finite=bool(torch.isfinite(G).all())))
    print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_lorenz.json", "w"), indent=1)
