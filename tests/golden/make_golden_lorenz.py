"""Golden vectors for the Lorenz 'pde'-type models, made with the REAL reference classes (agarbuno/ces at
/root/reference, ces/utils.py:124-447).  Run in the build container only:
    python tests/golden/make_golden_lorenz.py
Writes tests/golden/lorenz_cases.npz:
  l63 / l63log   short trajectories (T = 2, 201 samples) from lorenz63.solve (scipy odeint) and their statistics
  l96_*          right-hand sides of lorenz96 and its parameter-subset subclasses at a random state, and the
                 statistics of lorenz96 / lorenz96_hom applied to a given (random) trajectory array -- the parts of the
                 two-scale model that are deterministic functions.  Its trajectories are not stored: with the reference's
                 RK45 tolerances they differ from the true solution by O(1) after half a time unit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from ces import utils as ru  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(7)
    t = np.arange(0, 2.0 + 1e-9, 0.01)
    out["l63_t"] = t
    w0 = np.array([[1.0, 2.0, 20.0], [-5.0, -7.0, 25.0], [8.0, 1.0, 30.0]])
    args = np.array([[28.0, 8.0 / 3], [25.0, 2.0], [35.0, 3.0]])
    m = ru.lorenz63(l_window=1, freq=100)
    ml = ru.lorenz63_log(l_window=1, freq=100)
    out["l63_w0"], out["l63_args"] = w0, args
    out["l63_ws"] = np.stack([m.solve(w0[i], t, args=tuple(args[i])) for i in range(3)])
    out["l63_stats"] = np.stack([m.statistics(ws) for ws in out["l63_ws"]])
    out["l63log_ws"] = np.stack([ml.solve(w0[i], t, args=tuple(np.log(args[i]))) for i in range(3)])
    out["l63log_stats"] = np.stack([ml.statistics(ws) for ws in out["l63log_ws"]])
    # Lorenz 96: RHS of every variant at a random state
    ns, nf = 6, 4
    w = rng.standard_normal(ns * (nf + 1))
    out["l96_state"] = w
    for name, cls, vals in (("full", ru.lorenz96, (0.8, 9.0, np.log(8.0), 11.0)), ("Fc", ru.lorenz96Fc, (9.0, np.log(8.0))),
                            ("Fb", ru.lorenz96Fb, (9.0, 11.0)), ("hFb", ru.lorenz96hFb, (0.8, 9.0, 11.0)),
                            ("hcb", ru.lorenz96hcb, (0.8, np.log(8.0), 11.0))):
        mod = cls() if cls is not ru.lorenz96 else cls(n_slow=ns, n_fast=nf)
        mod.n_slow, mod.n_fast, mod.n_state = ns, nf, ns * (nf + 1)
        out["l96_args_" + name] = np.asarray(vals)
        out["l96_rhs_" + name] = np.asarray(mod(0.0, w, *vals))
    # statistics of a given trajectory
    mod = ru.lorenz96(n_slow=ns, n_fast=nf, l_window=1, freq=10, spinup=1)
    traj = rng.standard_normal((31, ns * (nf + 1)))
    out["l96_traj"] = traj
    out["l96_stats"] = mod.statistics(traj)
    hom = ru.lorenz96_hom()
    hom.n_slow, hom.n_fast, hom.n_state, hom.l_window, hom.freq, hom.spinup = 8, 4, 40, 1, 10, 1
    traj8 = rng.standard_normal((31, 40))
    out["l96hom_traj"] = traj8
    out["l96hom_stats"] = hom.statistics(traj8)
    hom.hom = False
    out["l96hom_stats_k7"] = hom.statistics(traj8)
    np.savez_compressed(os.path.join(HERE, "lorenz_cases.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
