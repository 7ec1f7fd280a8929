"""Golden vectors of the REAL reference (agarbuno/ces at /root/reference) for the non-default time_step modes:
'constant', 'mix' (before / after the re-solve threshold and after spin-up) and 'spectral', for eks_update and
eks_update_aldi.  Run in the build container only:
    python tests/golden/make_golden_timestep.py
Writes tests/golden/timestep_cases.npz (inputs once per case; per (rule, mode, history) the reference's Uk, hk, t).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import eks_oracle as eo, reference_loader as rl  # noqa: E402

CASES = [("small_dense", 6, 9, 40, True), ("J_less_than_k", 20, 33, 17, False), ("mid", 24, 50, 200, True)]
MODES = [("constant", None), ("constant", [0.4, 0.9]), ("mix", None), ("mix", [0.5, 1.7]), ("mix", [2.0, 5.0]),
         ("spectral", None), ("spectral", [0.3])]


def main():
    if not rl.available():
        raise SystemExit("reference not found at %s" % rl.REFERENCE_ROOT)
    out = {"names": np.array([c[0] for c in CASES]),
           "modes": np.array(["%s|%s" % (m, "" if th is None else ",".join(map(str, th))) for m, th in MODES])}
    for (name, d, k, J, dense) in CASES:
        pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense)
        for key in ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi"):
            out["%s/%s" % (name, key)] = pr[key]
        for rule in ("eks", "aldi"):
            for i, (mode, th) in enumerate(MODES):
                Uk, hk, m = rl.reference_step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"],
                                              pr["ustar"], pr["xi"], t_hist=th, time_step=mode, delta_t=0.03)
                out["%s/%s/%d/Uk" % (name, rule, i)] = Uk
                out["%s/%s/%d/hk_t" % (name, rule, i)] = np.array([hk, m["t"]])
    np.savez_compressed(os.path.join(HERE, "timestep_cases.npz"), **out)
    print("wrote", len(CASES), "cases x 2 rules x", len(MODES), "modes")


if __name__ == "__main__":
    main()
