"""The scenario of the reference's examples/scripts/darcy-flow.py on the B200 path.

Same flow as the reference script (problem set-up :9-36, run_neks :40-93, the ensemble-size sweep :97-105) with the
two import lines swapped; the MATLAB engine calls (`model.start`, `model.set_rnd_seed`) are kept and are no-ops here.

    python examples/darcy_flow.py [--nmesh 16] [--T 200] [--t-tol 5] [--rng numpy|device] [--save-online]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ces_b200.calibrate import *      # noqa: F401,F403   (reference: from ces.calibrate import *)
import ces_b200.darcy as darcy         # (reference: import ces.darcy as darcy)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=16)
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--t-tol", type=float, default=5.0)
    ap.add_argument("--rng", default="numpy")
    ap.add_argument("--save-online", action="store_true")
    ap.add_argument("--sizes", default="")
    args = ap.parse_args()

    model = darcy.model(Nmesh=args.nmesh)
    model.start(mpath=r'./mfiles')
    model.set_rnd_seed()
    model.set_initial()
    model.n_obs = 50
    model.model_name = 'darcy-flow'

    # Forward model G(u) in D: KL coefficients -> pressure on the grid
    U = model(model.ustar, full_solution=True)

    # Random observation locations, sampled proportionally to the pressure
    xs, ys = np.meshgrid(np.linspace(0, 1, int(model.Nmesh)), np.linspace(0, 1, int(model.Nmesh)))
    np.random.seed(1)
    grid_pts = np.vstack((xs.flatten(), ys.flatten()))
    model.obs_index = np.random.choice(int(model.p), model.n_obs, replace=False, p=U / U.sum())
    model.obs_locs = grid_pts[:, model.obs_index]

    y_obs = model(model.ustar)
    gamma = 0.005
    Gamma = gamma ** 2 * np.identity(model.n_obs)
    y_obs = y_obs + 1.0 * gamma * np.random.normal(0, 1, model.n_obs)
    print('Remember, the number of params is: %s' % (len(model.ustar)))
    Jnoise = np.linalg.cholesky(Gamma)

    def run_neks(J, model, **kwargs):
        eks = sampling(p=model.p, n_obs=model.n_obs, J=J)
        eks.ustar = model.ustar.reshape(model.p, -1)
        eks.T = args.T
        eks.mu = 0.0 * np.ones((model.p,)).reshape(model.p, -1)
        eks.sigma = 100. * np.identity(model.p)
        eks.parallel = False
        eks.mute_bar = True
        eks.nexp = kwargs.get('nexp', '')
        eks.directory = './'
        np.random.seed(kwargs.get('nexp', 1))
        U0 = 10 * np.random.normal(0, 1, [eks.p, J])
        eks.run(y_obs, U0, model, Gamma, Jnoise, save_online=args.save_online, t_tol=args.t_tol, rng=args.rng)
        return eks

    if args.sizes:
        Js = [int(s) for s in args.sizes.split(",")]
    else:
        Js = [int(model.p / 15), int(model.p / 5), int(model.p / 2), int(model.p + 2), int(2 * model.p), int(3 * model.p)]
    np.random.seed(1)
    neks = {}
    for J in Js:
        t0 = time.perf_counter()
        eks = run_neks(J, model, nexp=0)
        dt = time.perf_counter() - t0
        neks['eks-' + str(J).zfill(3)] = [eks]
        n = len(eks.metrics['t'])
        print("J=%4d  iterations=%3d  t=%.3f  bias=%.4g -> %.4g  bias-data=%.4g -> %.4g  %.1f ms/iteration (CG its %d)"
              % (J, n, eks.metrics['t'][-1], eks.metrics['bias'][0], eks.metrics['bias'][-1],
                 eks.metrics['bias-data'][0], eks.metrics['bias-data'][-1], 1e3 * dt / max(n, 1), model.last_iterations))
    model.stop()
    return neks


if __name__ == "__main__":
    main()
