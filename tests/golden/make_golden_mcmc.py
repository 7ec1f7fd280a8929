"""Golden vectors of the REAL reference's MCMC.model_mh (ces/sample.py:121-196) on the map models.

Run in the build container only:  python tests/golden/make_golden_mcmc.py   -> tests/golden/mcmc_cases.npz
The reference file is exec'd unmodified with stand-ins for its unavailable imports (oracle/reference_loader.load_sample);
the forward models and enka.G are the reference's own (ces/utils.py, ces/calibrate.py).  Every case stores its inputs, the
sample chain, the acceptance rate and the state of numpy's global generator after the run.
"""
import os
import sys

import numpy as np
from scipy.stats import multivariate_normal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_loader as rl  # noqa: E402


def cases():
    rs = np.random.RandomState(5)
    A = np.ones((10, 2))
    A[:, 1] = 2 * rs.normal(0, 1, 10)
    ustar = np.array([-1.0, 2.0])
    lin = dict(kind="lineal", A=A, y=A @ ustar + np.sqrt(0.1) * rs.normal(0, 1, 10), Gamma=0.1 * np.eye(10),
               Ustar=ustar[:, None] + 0.3 * rs.normal(size=(2, 50)), prior_mean=np.zeros(2), prior_cov=100.0 * np.eye(2))
    A3 = rs.normal(size=(7, 3))
    Q = rs.normal(size=(7, 7))
    lin3 = dict(kind="lineal", A=A3, y=A3 @ np.array([0.5, -0.2, 1.0]), Gamma=0.05 * (np.eye(7) + 0.3 * Q @ Q.T / 7),
                Ustar=np.array([[0.5], [-0.2], [1.0]]) + 0.2 * rs.normal(size=(3, 40)), prior_mean=np.array([0.1, 0.0, -0.1]),
                prior_cov=4.0 * np.eye(3) + 0.5)
    ell = dict(kind="elliptic", y=np.array([27.45194112300398, 79.70194112300398]), Gamma=0.01 * np.eye(2),
               Ustar=np.array([[-2.65], [104.5]]) + np.array([[0.05], [0.5]]) * rs.normal(size=(2, 30)),
               prior_mean=np.array([-2.0, 100.0]), prior_cov=np.diag([1.0, 100.0]))
    ban = dict(kind="banana", y=np.array([0.4, 0.3]), Gamma=(0.55 ** 2) * np.array([[1.0, 0.9], [0.9, 1.0]]),
               Ustar=np.array([[0.4], [1.0]]) + 0.5 * rs.normal(size=(2, 30)), prior_mean=np.zeros(2), prior_cov=np.eye(2) * 9.0)
    lin_pcn = dict(lin, prior_mean=np.zeros(2), prior_cov=np.array([[1.0, 0.3], [0.3, 4.0]]))
    return [("lineal_rw", lin, {}, 1.0, True, 400), ("lineal_pcn", lin_pcn, {"update": "pCN", "beta": 0.001}, 1.0, True, 300),
            ("lineal3_dense", lin3, {}, 0.8, True, 300), ("lineal_noscale", lin, {}, 0.05, False, 300),
            ("elliptic_rw", ell, {}, 1.5, True, 300), ("banana_rw", ban, {}, 1.0, True, 300)]


def main():
    if not rl.available():
        raise SystemExit("reference not found at %s" % rl.REFERENCE_ROOT)
    cal, utils, sample = rl.load_calibrate(), rl.load_utils(), rl.load_sample()
    out, names = {}, []
    for name, c, kw, delta, scaling, n in cases():
        p, J = c["Ustar"].shape
        k = c["y"].shape[0]
        enka = cal.sampling(p=p, n_obs=k, J=J)
        enka.Ustar = c["Ustar"]
        model = {"lineal": lambda: utils.lineal(c["A"]), "elliptic": utils.elliptic, "banana": utils.banana}[c["kind"]]()
        prior = multivariate_normal(c["prior_mean"], c["prior_cov"])
        mc = sample.MCMC()
        mc.mute_bar = True
        mc.y_obs = c["y"]
        np.random.seed(13)
        mc.model_mh(model, n, prior, enka, c["Gamma"], delta=delta, enka_scaling=scaling, **kw)
        first, first_accept = mc.samples.copy(), float(mc.accept)
        state1 = np.random.get_state()
        mc.model_mh(model, 50, prior, enka, c["Gamma"], delta=delta, enka_scaling=scaling, **kw)      # resume
        state2 = np.random.get_state()
        for key in ("y", "Gamma", "Ustar", "prior_mean", "prior_cov"):
            out["%s/%s" % (name, key)] = c[key]
        if "A" in c:
            out["%s/A" % name] = c["A"]
        out["%s/kind" % name] = np.array(c["kind"])
        out["%s/update" % name] = np.array(kw.get("update", ""))
        out["%s/beta" % name] = np.float64(kw.get("beta", 0.5))
        out["%s/delta" % name] = np.float64(delta)
        out["%s/scaling" % name] = np.bool_(scaling)
        out["%s/n" % name] = np.int64(n)
        out["%s/samples" % name] = first
        out["%s/accept" % name] = np.float64(first_accept)
        out["%s/samples_resumed" % name] = mc.samples
        out["%s/accept_resumed" % name] = np.float64(mc.accept)
        for tag, st in (("state1", state1), ("state2", state2)):
            out["%s/%s_key" % (name, tag)] = st[1]
            out["%s/%s_pos" % (name, tag)] = np.array([st[2], st[3]])
            out["%s/%s_gauss" % (name, tag)] = np.float64(st[4])
        names.append(name)
        print(name, first.shape, "accept %.3f, resumed run %.3f" % (first_accept, mc.accept))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mcmc_cases.npz"), **out)


if __name__ == "__main__":
    main()
