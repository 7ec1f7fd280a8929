"""MCMC.model_mh on the device (ces_b200/sample.py, csrc/mcmc.cu) against chains of the REAL reference
(ces/sample.py:121-196; tests/golden/make_golden_mcmc.py): same samples, same acceptance rate, same state of numpy's
global generator afterwards -- the kernel consumes the MT19937 stream the reference's np.random.normal / uniform calls
would have consumed -- for random walk and pCN, enka-scaled and fixed proposals, dense Gamma, the resume path, all four
map models."""
import os

import numpy as np
import pytest
from scipy.stats import multivariate_normal

pytestmark = pytest.mark.gpu

from ces_b200 import calibrate, sample as csample, utils as cutils  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcmc_cases.npz"))


def _setup(name):
    c = {key: G["%s/%s" % (name, key)] for key in ("y", "Gamma", "Ustar", "prior_mean", "prior_cov")}
    kind = str(G["%s/kind" % name])
    p, J = c["Ustar"].shape
    enka = calibrate.sampling(p, c["y"].shape[0], J)
    enka.Ustar = c["Ustar"]
    model = {"lineal": lambda: cutils.lineal(G["%s/A" % name]), "elliptic": cutils.elliptic, "banana": cutils.banana}[kind]()
    prior = multivariate_normal(c["prior_mean"], c["prior_cov"])
    kw = {}
    if str(G["%s/update" % name]):
        kw = {"update": str(G["%s/update" % name]), "beta": float(G["%s/beta" % name])}
    mc = csample.MCMC()
    mc.y_obs = c["y"]
    return mc, model, prior, enka, c["Gamma"], float(G["%s/delta" % name]), bool(G["%s/scaling" % name]), kw, int(G["%s/n" % name])


def _state_equal(name, tag):
    st = np.random.get_state()
    return (np.array_equal(st[1], G["%s/%s_key" % (name, tag)]) and [st[2], st[3]] == list(G["%s/%s_pos" % (name, tag)])
            and (st[3] == 0 or abs(st[4] - float(G["%s/%s_gauss" % (name, tag)])) <= 1e-14 * max(1.0, abs(st[4]))))


@pytest.mark.parametrize("name", [str(n) for n in G["names"]])
def test_model_mh_reproduces_the_reference_chain(name):
    mc, model, prior, enka, Gamma, delta, scaling, kw, n = _setup(name)
    np.random.seed(13)
    mc.model_mh(model, n, prior, enka, Gamma, delta=delta, enka_scaling=scaling, **kw)
    want = G["%s/samples" % name]
    assert mc.samples.shape == want.shape == (enka.p, n + 1)
    scale = np.abs(want).max()
    assert np.abs(mc.samples - want).max() / scale < 1e-10, name          # every accept decision and every state
    assert abs(mc.accept - float(G["%s/accept" % name])) < 1e-12
    assert _state_equal(name, "state1")                                       # the generator is where the reference leaves it
    # second call on the same object: continues from the last sample (:156-164)
    mc.model_mh(model, 50, prior, enka, Gamma, delta=delta, enka_scaling=scaling, **kw)
    want2 = G["%s/samples_resumed" % name]
    assert mc.samples.shape == want2.shape == (enka.p, n + 51)
    assert np.abs(mc.samples - want2).max() / scale < 1e-10
    assert abs(mc.accept - float(G["%s/accept_resumed" % name])) < 1e-12 and _state_equal(name, "state2")


def test_many_chains_sample_the_linear_gaussian_posterior():
    """n_chains independent device chains (Philox noise) beside the reference's one: pooled, they reproduce the analytic
    posterior of the linear-Gaussian problem (mean and covariance)."""
    name = "lineal_rw"
    mc, model, prior, enka, Gamma, delta, scaling, kw, n = _setup(name)
    A, y = G["%s/A" % name], G["%s/y" % name]
    post_cov = np.linalg.inv(A.T @ np.linalg.solve(Gamma, A) + np.linalg.inv(prior.cov))
    post_mean = post_cov @ (A.T @ np.linalg.solve(Gamma, y) + np.linalg.solve(prior.cov, prior.mean))
    np.random.seed(13)
    mc.model_mh(model, 3000, prior, enka, Gamma, delta=0.15, enka_scaling=False, n_chains=256, seed=5)
    ch = mc.samples_chains
    assert ch.shape == (256, 2, 3001) and mc.accept_chains.shape == (256,) and np.all(mc.accept_chains > 0.05)
    pooled = ch[:, :, 1000:].transpose(1, 0, 2).reshape(2, -1)
    assert np.abs(pooled.mean(axis=1) - post_mean).max() < 0.02
    ratio = np.cov(pooled) / post_cov
    assert np.all(np.abs(np.diag(ratio) - 1.0) < 0.15)
    # chain 0 is still the reference-stream chain; the others differ from it and from each other
    assert not np.array_equal(ch[1], ch[2]) and np.array_equal(ch[0], mc.samples)


def test_model_mh_refuses_what_it_does_not_cover():
    mc, model, prior, enka, Gamma, delta, scaling, kw, n = _setup("lineal_rw")

    class Host(object):
        type = "map"

        def __call__(self, theta):
            return theta

    with pytest.raises(NotImplementedError):
        mc.model_mh(Host(), 10, prior, enka, Gamma)
    with pytest.raises(NotImplementedError):
        mc.gp_mh(enka, 10, prior)


def test_calibrate_then_sample_example():
    """examples/linear_ces.py: the reference's Calibrate -> Sample hand-off (linear.ipynb scenario) end to end on the device:
    sampling.run finds the posterior region, MCMC.model_mh scaled by eks.Ustar samples it; the pooled chains reproduce the
    analytic posterior (the notebook prints mean [-1.0367 2.0870], cov [[0.01007 0.00034] [0.00034 0.00176]])."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("linear_ces_example", os.path.join(root, "examples", "linear_ces.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--J", "100", "--T", "400", "--n-mcmc", "6000", "--chains", "64"])
    assert np.allclose(out["posterior_mean"], [-1.03673079, 2.08697021], atol=2e-3)
    assert np.abs(out["eks_mean"] - out["posterior_mean"]).max() < 0.08              # one ensemble of 100 particles
    assert np.abs(out["mcmc_mean"] - out["posterior_mean"]).max() < 0.01
    ratio = np.diag(out["mcmc_cov"]) / np.diag(out["posterior_cov"])
    assert np.all(np.abs(ratio - 1.0) < 0.15) and 0.05 < out["accept"] < 0.9
    assert out["mcmc"].samples.shape == (2, 6001) and out["mcmc"].samples_chains.shape == (64, 2, 6001)


@pytest.mark.parametrize("p,k,kind,update", [(1, 1, "lineal", None), (2, 10, "lineal_log", None), (5, 64, "lineal", "pCN"),
                                             (32, 7, "lineal", None), (32, 64, "lineal", None), (17, 33, "lineal_log", "pCN"),
                                             (3, 40, "lineal", None)])
def test_model_mh_matches_the_oracle_on_fresh_problems(p, k, kind, update):
    """Device chain against oracle/mcmc_oracle.py (itself pinned to the real reference's chains) on seeded problems the
    golden file does not hold: the limits of the kernel (p = 32, k = 64), a single parameter / observation, dense Gamma and
    prior covariance, lineal_log, pCN -- samples, acceptance and the generator state."""
    from oracle import forward_oracle as fo, mcmc_oracle as mo

    rs = np.random.RandomState(100 * p + k)
    A = rs.normal(size=(k, p)) / np.sqrt(p) * (0.3 if kind == "lineal_log" else 1.0)
    truth = 0.3 * rs.normal(size=p)
    fwd = (lambda th: fo.lineal(A, np.asarray(th).reshape(-1, 1))[:, 0]) if kind == "lineal" else \
          (lambda th: fo.lineal_log(A, np.asarray(th).reshape(-1, 1))[:, 0])
    y = fwd(truth) + 0.05 * rs.normal(size=k)
    Q = rs.normal(size=(k, k))
    Gamma = 0.01 * (np.eye(k) + 0.3 * Q @ Q.T / k)
    S = rs.normal(size=(p, p))
    prior = multivariate_normal(0.1 * rs.normal(size=p) if update is None else np.zeros(p), np.eye(p) + 0.2 * S @ S.T / p)
    J = 3 * p + 8
    Ustar = truth[:, None] + 0.05 * rs.normal(size=(p, J))
    kw = {"delta": 0.6, "enka_scaling": True}
    if update:
        kw.update(update=update, beta=0.002)
    np.random.seed(77)
    want, want_acc = mo.model_mh(fwd, 200, prior, Ustar, y, Gamma, **kw)
    state = np.random.get_state()
    enka = calibrate.sampling(p, k, J)
    enka.Ustar = Ustar
    mc = csample.MCMC()
    mc.y_obs = y
    model = cutils.lineal(A) if kind == "lineal" else cutils.lineal_log(A)
    np.random.seed(77)
    mc.model_mh(model, 200, prior, enka, Gamma, **kw)
    assert mc.samples.shape == want.shape == (p, 201)
    assert np.abs(mc.samples - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-300), (p, k, kind, update)
    assert abs(mc.accept - want_acc) < 1e-12 and 0.0 < want_acc < 1.0
    got = np.random.get_state()
    assert np.array_equal(got[1], state[1]) and got[2] == state[2] and got[3] == state[3]
