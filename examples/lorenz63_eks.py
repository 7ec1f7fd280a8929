"""The scenario of the reference's examples/notebooks/lorenz63.ipynb (cells 8-15, 19) on the B200 path.

Parameters (log r, log b) of the Lorenz 63 system are calibrated from time-averaged moments: data generation from a long
trajectory, noise covariance from the rolling-window statistics, `sampling.run` with the state carried over between
iterations, and the notebook's direct use of `G_pde_ens` on a parameter grid.  Only the two import lines differ from the
notebook (`from ces.utils import *`, `from ces.calibrate import *`).

    python examples/lorenz63_eks.py [--J 100] [--T 10] [--T-eks 20] [--grid 24]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ces_b200.utils import *          # noqa: F401,F403   (reference: from ces.utils import *)
from ces_b200.calibrate import *      # noqa: F401,F403   (reference: from ces.calibrate import *)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--J", type=int, default=100)
    ap.add_argument("--T", type=int, default=10, help="iterations of the sampler (notebook: 50)")
    ap.add_argument("--T-data", type=float, default=360.0, help="length of the data trajectory (notebook: 360)")
    ap.add_argument("--T-eks", type=float, default=20.0, help="integration time per forward evaluation (notebook: 60)")
    ap.add_argument("--grid", type=int, default=24, help="side of the (log r, log b) grid for the misfit surface (notebook: 60)")
    args = ap.parse_args()

    T, dt = args.T_data, 100.0                                  # final time, samples per time unit (cell 8)
    t = np.linspace(0, T, int(T * dt) + 1)
    T_roll, T_spinup = 10.0, 30.0
    w0 = (1.0, 1.0, 1.0)

    model = lorenz63_log()                                      # cell 9
    model.l_window = T_roll
    model.freq = dt
    ws = model.solve(w0, t)

    n_obs, p = 9, 2                                             # cell 10
    ustar = np.array([[np.log(28.0)], [np.log(8.0 / 3)]])
    xs, ys, zs = ws[:, 0], ws[:, 1], ws[:, 2]
    mom = np.asarray([xs, ys, zs, xs ** 2, ys ** 2, zs ** 2, xs * ys, xs * zs, ys * zs])
    wt = np.asarray([xs[-1], ys[-1], zs[-1]])
    win = int(T_roll * dt)
    csum = np.cumsum(np.insert(mom, 0, 0.0, axis=1), axis=1)
    gs = np.full(mom.shape, np.nan)
    gs[:, win - 1:] = (csum[:, win:] - csum[:, :-win]) / win    # pd.Series(k).rolling(window).mean()
    Gamma = np.cov(gs[:, t > T_spinup])
    y_obs = gs[:, t > T_spinup].mean(axis=1)

    Jnoise = np.linalg.cholesky(Gamma)                          # cell 13
    t_eks = np.linspace(0, args.T_eks, int(args.T_eks * dt) + 1)
    enki = sampling(p=p, n_obs=n_obs, J=args.J)
    enki.ustar = ustar
    enki.T = args.T
    enki.mu = np.array([3.3, 1.2]).reshape(2, -1)
    enki.sigma = np.diag([0.15 ** 2, 0.5 ** 2])
    enki.parallel = True
    enki.mute_bar = True
    np.random.seed(2016)
    U0 = enki.mu + enki.sigma ** 0.5 @ np.random.normal(0, 1, [enki.p, enki.J])

    t0 = time.perf_counter()
    enki.run(y_obs, U0, model, Gamma, Jnoise, wt=wt, t=t_eks)   # cell 15
    dt_run = time.perf_counter() - t0
    n = len(enki.metrics['t'])
    print("EKS: %d iterations in %.2f s, t = %.3f, ensemble mean (r, b) = (%.3f, %.3f), truth (28, 2.667)"
          % (n, dt_run, enki.metrics['t'][-1], np.exp(enki.Ustar[0]).mean(), np.exp(enki.Ustar[1]).mean()))

    # cell 19: the data misfit on a grid of parameters, initial states resampled from the data trajectory
    grid = args.grid
    rs, bs = np.meshgrid(np.linspace(np.log(24.0), np.log(33.0), grid), np.linspace(np.log(1.8), np.log(3.6), grid))
    starts = np.asarray([xs, ys, zs])[:, np.random.choice(6001, grid ** 2, replace=True)]
    t0 = time.perf_counter()
    Gs = enki.G_pde_ens(np.vstack([np.array([rs.flatten(), bs.flatten()]), starts]), model, t_eks)
    Phi = ((Gs[:n_obs] - y_obs[:, None]) * np.linalg.solve(2 * Gamma, Gs[:n_obs] - y_obs[:, None])).sum(axis=0)
    k = int(np.argmin(Phi))
    print("misfit surface on %d x %d parameters in %.2f s; minimum at (r, b) = (%.2f, %.2f)"
          % (grid, grid, time.perf_counter() - t0, np.exp(rs.flatten()[k]), np.exp(bs.flatten()[k])))
    return enki, Phi


if __name__ == "__main__":
    main()
