"""Golden vectors of the REAL reference for a resumed sampling.run (ces/calibrate.py:307-315, 329-339): the second call of
run() on the same object continues Uall / Gall / metrics (cumulative pseudo-time included) instead of starting over.

Run in the build container only:  python tests/golden/make_golden_resume.py   -> tests/golden/resume_case.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_loader as rl  # noqa: E402


def problem():
    rs = np.random.RandomState(11)
    d, k, J = 3, 7, 40
    A = rs.normal(size=(k, d))
    ustar = rs.normal(size=(d, 1))
    return dict(A=A, ustar=ustar, y=A @ ustar[:, 0] + 0.1 * rs.normal(size=k), Gamma=0.01 * np.eye(k), mu=np.zeros((d, 1)),
                Sigma0=25.0 * np.eye(d), U0=2.0 * rs.normal(size=(d, J)))


def main():
    if not rl.available():
        raise SystemExit("reference not found at %s" % rl.REFERENCE_ROOT)
    cal, utils = rl.load_calibrate(), rl.load_utils()
    pr = problem()
    d, J = pr["U0"].shape
    eks = cal.sampling(p=d, n_obs=pr["A"].shape[0], J=J)
    eks.ustar, eks.mu, eks.sigma, eks.T = pr["ustar"], pr["mu"], pr["Sigma0"], 3
    model = utils.lineal(pr["A"])
    np.random.seed(21)
    eks.run(pr["y"], pr["U0"], model, pr["Gamma"], None, t_tol=1e9)
    first = dict(Uall=np.array(eks.Uall), t=list(eks.metrics["t"]))
    eks.T = 2
    eks.run(pr["y"], eks.Ustar, model, pr["Gamma"], None, t_tol=1e9)          # resumes: same object, same RNG stream
    out = {key: pr[key] for key in pr}
    out.update(first_Uall=first["Uall"], first_t=np.array(first["t"]), Uall=np.array(eks.Uall), Gall=np.array(eks.Gall),
               Ustar=eks.Ustar, Gstar=eks.Gstar,
               metrics=np.array([eks.metrics[key] for key in ("self-bias", "bias", "self-bias-data", "bias-data", "t")]))
    # banana: the reference's model draws two normals per evaluation even without noise (ces/utils.py:122), so the update
    # noise of a seeded run is interleaved with them; a drop-in must consume the same stream
    ban = utils.banana()
    rs = np.random.RandomState(2)
    Ub = np.array([[0.4], [1.0]]) + 0.5 * rs.normal(size=(2, 30))
    eb = cal.sampling(p=2, n_obs=2, J=30)
    eb.ustar, eb.mu, eb.sigma, eb.T = np.array([[0.4], [1.0]]), np.zeros((2, 1)), 9.0 * np.eye(2), 4
    np.random.seed(8)
    eb.run(np.array([0.4, 0.3]), Ub, ban, ban.Gamma, None, t_tol=1e9)
    out.update(banana_U0=Ub, banana_Gamma=ban.Gamma, banana_Uall=np.array(eb.Uall), banana_Gall=np.array(eb.Gall),
               banana_t=np.array(eb.metrics["t"]), banana_next_normal=np.random.normal(0, 1, 3))
    np.savez_compressed(os.path.join(HERE, "resume_case.npz"), **out)
    print("Uall", out["Uall"].shape, "t", out["metrics"][4])


if __name__ == "__main__":
    main()
