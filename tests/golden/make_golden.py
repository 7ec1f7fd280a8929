"""Generate the committed golden vectors from the REAL reference (agarbuno/ces at /root/reference).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/step_cases.npz and tests/golden/forward_cases.npz.  Each step case stores the
inputs (y, U, G, Gamma, mu, sigma, ustar, xi) and the reference outputs (Uk, hk, the four metrics, t)
of sampling.eks_update / eks_update_aldi / eks_update_aldi_constant run through
oracle/reference_loader.py (tab-expanded in memory, noise injected through np.random.normal).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import eks_oracle as eo, reference_loader as rl  # noqa: E402

# (name, d, k, J, dense Gamma, dense Sigma0, t history)
CASES = [
    ("cfg1_linear", 2, 10, 100, False, False, None),
    ("ragged_small", 3, 5, 33, True, True, None),
    ("J_less_than_d", 40, 30, 17, True, False, [0.4]),
    ("cfg2_shape", 64, 50, 256, False, False, [0.3, 1.7]),
    ("dense_both", 24, 36, 130, True, True, None),
]
RULES = ("eks", "aldi", "aldi_constant")


def build_case(d, k, J, dense_g, dense_s):
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense_g)
    rng = np.random.default_rng(5)
    if dense_s:
        S = rng.standard_normal((d, d))
        pr["Sigma0"] = 50 * np.eye(d) + S @ S.T
        pr["mu"] = rng.standard_normal((d, 1))
    return pr


def main():
    if not rl.available():
        raise SystemExit("reference not found at %s" % rl.REFERENCE_ROOT)
    out = {}
    names = []
    for (name, d, k, J, dg, ds, th) in CASES:
        pr = build_case(d, k, J, dg, ds)
        for key in ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi"):
            out["%s/%s" % (name, key)] = pr[key]
        out["%s/t_hist" % name] = np.asarray(th if th else [], dtype=float)
        for rule in RULES:
            Uk, hk, m = rl.reference_step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"],
                                          pr["ustar"], pr["xi"], t_hist=th)
            out["%s/%s/Uk" % (name, rule)] = Uk
            out["%s/%s/hk" % (name, rule)] = np.float64(hk)
            out["%s/%s/metrics" % (name, rule)] = np.array([m["self-bias"], m["bias"], m["self-bias-data"],
                                                           m["bias-data"], m["t"]])
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "step_cases.npz"), **out)

    # forward maps through the reference's own enka.G_ens (per-particle Python loop)
    utils = rl.load_utils()
    cal = rl.load_calibrate()
    rng = np.random.default_rng(11)
    fw = {}
    A = rng.standard_normal((10, 4))
    U = rng.standard_normal((4, 37))
    e = cal.enka(4, 10, 37)
    fw["lineal/A"], fw["lineal/U"] = A, U
    fw["lineal/b"] = np.float64(0.7)
    fw["lineal/G"] = e.G_ens(U, utils.lineal(A, b=0.7))
    fw["lineal_log/G"] = e.G_ens(U, utils.lineal_log(A))
    U2 = rng.standard_normal((2, 53)) * np.array([[1.0], [50.0]])
    e2 = cal.enka(2, 2, 53)
    fw["map2/U"] = U2
    fw["elliptic/G"] = e2.G_ens(U2, utils.elliptic())
    fw["banana/G"] = e2.G_ens(U2, utils.banana(a=1.3, b=0.4))
    # problem constants printed in the notebooks (examples/notebooks/elliptic.ipynb:72, :112)
    fw["elliptic/y_obs_notebook"] = np.array([27.45194112300398, 79.70194112300398])
    fw["elliptic/ustar_notebook"] = np.array([-2.65, 104.5])
    fw["elliptic/G_at_ustar"] = np.asarray(utils.elliptic()(np.array([-2.65, 104.5])), dtype=float)
    np.savez_compressed(os.path.join(HERE, "forward_cases.npz"), **fw)
    print("wrote", len(names), "step cases x", len(RULES), "rules and the forward cases")


if __name__ == "__main__":
    main()
