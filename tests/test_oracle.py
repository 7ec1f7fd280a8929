"""The oracle against the reference's own material (CPU, no GPU needed):
 * golden vectors generated from the real reference (tests/golden/make_golden.py),
 * the real reference itself when /root/reference is present (build container only),
 * the analytic linear-Gaussian posterior printed in examples/notebooks/linear.ipynb:692-697,
 * the elliptic problem constants of examples/notebooks/elliptic.ipynb:72,112.
"""
import numpy as np
import pytest

from oracle import eks_oracle as eo, forward_oracle as fo, reference_loader as rl

RULES = ("eks", "aldi", "aldi_constant")
TOL = 1e-12   # oracle vs reference: same LAPACK, only expression order differs


def _case(golden, name):
    keys = ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi", "t_hist")
    return {k: golden["%s/%s" % (name, k)] for k in keys}


def test_oracle_matches_golden(golden_steps):
    for name in golden_steps["names"]:
        c = _case(golden_steps, name)
        t_last = float(c["t_hist"][-1]) if len(c["t_hist"]) else None
        for rule in RULES:
            o = eo.step(rule, c["y"], c["U0"], c["G"], c["Gamma"], c["mu"], c["Sigma0"], c["ustar"], c["xi"],
                        t_last=t_last)
            Uk = golden_steps["%s/%s/Uk" % (name, rule)]
            m = golden_steps["%s/%s/metrics" % (name, rule)]
            assert np.abs(o["Uk"] - Uk).max() / np.abs(Uk).max() < TOL, (name, rule)
            got = [o["metrics"][q] for q in ("self-bias", "bias", "self-bias-data", "bias-data")] + [o["t"]]
            assert np.allclose(got, m, rtol=1e-12, atol=0), (name, rule)
            if t_last is None or rule != "aldi_constant":
                assert abs(o["hk"] - float(golden_steps["%s/%s/hk" % (name, rule)])) <= 1e-12 * o["hk"]


def test_as_written_metrics_equal_cheap(golden_steps):
    c = _case(golden_steps, "ragged_small")
    E, R, W, D = eo.interaction(c["G"], c["y"], c["Gamma"])
    a = eo.metrics_as_written(c["U0"], c["ustar"], E, R, c["Gamma"])
    b = eo.metrics_cheap(c["U0"], c["ustar"], E, R, c["Gamma"])
    for key in a:
        assert abs(a[key] - b[key]) <= 1e-12 * abs(a[key])


@pytest.mark.skipif(not rl.available(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("ts", [None, "constant", "mix", "spectral"])
def test_oracle_matches_live_reference(ts):
    d, k, J = 6, 9, 40
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=True)
    for th in (None, [0.5, 1.5], [2.0, 5.0]):
        for rule in RULES:
            if rule == "aldi_constant" and ts:
                continue
            kw = {} if ts is None else {"time_step": ts}
            Uk, hk, m = rl.reference_step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"],
                                          pr["ustar"], pr["xi"], t_hist=th, **kw)
            o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"],
                        time_step=ts, t_last=None if th is None else th[-1])
            assert np.abs(o["Uk"] - Uk).max() / np.abs(Uk).max() < TOL
            assert abs(o["t"] - m["t"]) <= 1e-12 * abs(m["t"])


@pytest.mark.skipif(not rl.available(), reason="reference checkout not present (GPU box)")
def test_reference_run_loop_matches_oracle_loop():
    """sampling.run of the reference (forward -> update -> t_tol stop) against the oracle driven in a loop
    with the same numpy random stream."""
    utils = rl.load_utils()
    cal = rl.load_calibrate()
    np.random.seed(1)
    A = np.random.normal(size=(10, 2))
    ustar = np.array([[-1.0], [2.0]])
    Gamma = 0.1 * np.eye(10)
    y = A @ ustar[:, 0]
    J, T = 30, 12
    U0 = np.random.normal(0, 1, [2, J])
    eks = cal.sampling(p=2, n_obs=10, J=J)
    eks.ustar, eks.mu, eks.sigma, eks.T = ustar, np.zeros((2, 1)), 100.0 * np.eye(2), T
    np.random.seed(3)
    eks.run(y, U0, utils.lineal(A), Gamma, np.linalg.cholesky(Gamma), t_tol=1e9)
    np.random.seed(3)
    U, t = U0, None
    for _ in range(T):
        xi = np.random.normal(0, 1, [2, J])
        o = eo.step("aldi", y, U, fo.lineal(A, U), Gamma, eks.mu, eks.sigma, ustar, xi, t_last=t)
        U, t = o["Uk"], o["t"]
    assert np.abs(U - eks.Ustar).max() / np.abs(eks.Ustar).max() < 1e-9
    assert abs(t - eks.metrics["t"][-1]) < 1e-9 * t


def test_forward_oracle_matches_golden(golden_forward):
    g = golden_forward
    assert np.allclose(fo.lineal(g["lineal/A"], g["lineal/U"], float(g["lineal/b"])), g["lineal/G"], rtol=1e-13, atol=1e-13)
    assert np.allclose(fo.lineal_log(g["lineal/A"], g["lineal/U"]), g["lineal_log/G"], rtol=1e-13, atol=1e-13)
    assert np.allclose(fo.elliptic(g["map2/U"]), g["elliptic/G"], rtol=1e-13, atol=1e-13)
    assert np.allclose(fo.banana(g["map2/U"], a=1.3, b=0.4), g["banana/G"], rtol=1e-13, atol=1e-13)


def test_elliptic_notebook_constants(golden_forward):
    """examples/notebooks/elliptic.ipynb:72 prints y_obs for ustar = (-2.65, 104.5) (:112)."""
    g = golden_forward
    got = fo.elliptic(g["elliptic/ustar_notebook"].reshape(2, 1))[:, 0]
    assert np.allclose(got, g["elliptic/G_at_ustar"], rtol=1e-14)
    assert np.allclose(got, g["elliptic/y_obs_notebook"], rtol=1e-12)


def test_linear_gaussian_posterior_known_answer():
    """Long-run EKS on the linear.ipynb problem converges to the analytic posterior printed there
    (mean [-1.03673079 2.08697021], cov [[0.01006653 0.00034247],[0.00034247 0.00176275]],
    examples/notebooks/linear.ipynb:692-697): problem np.random.seed(1), A = [1, 2 N(0,1)] (10 x 2),
    u* = (-1, 2), noise 0.1, prior N(0, 10^2 I) (cell 4 / cell 11)."""
    np.random.seed(1)
    A = np.ones((10, 2))
    A[:, 1] = 2 * np.random.normal(0, 1, 10)
    ustar = np.array([-1.0, 2.0])
    noise = 0.1
    y = A @ ustar + np.sqrt(noise) * np.random.normal(0, 1, 10)
    Gamma = noise * np.eye(10)
    sigma2 = 100.0
    post_cov = np.linalg.inv(A.T @ A / noise + np.eye(2) / sigma2)
    post_mean = post_cov @ (A.T @ y / noise)
    # the notebook's printed analytic posterior (it is computed there without the 1/sigma^2 prior term,
    # hence agreement to ~1e-4 relative only)
    assert np.allclose(post_mean, [-1.03673079, 2.08697021], atol=2e-3)
    assert np.allclose(post_cov, [[0.01006653, 0.00034247], [0.00034247, 0.00176275]], rtol=2e-3, atol=1e-6)
    J = 100
    U = np.random.normal(0, 1, [2, J]) * 3.0
    t = None
    means, covs = [], []
    for it in range(1200):
        xi = np.random.normal(0, 1, [2, J])
        o = eo.step("aldi", y, U, A @ U, Gamma, np.zeros((2, 1)), sigma2 * np.eye(2), ustar.reshape(2, 1), xi, t_last=t)
        U, t = o["Uk"], o["t"]
        if it >= 400:
            means.append(U.mean(axis=1))
            covs.append(np.cov(U))
    m, c = np.mean(means, axis=0), np.mean(covs, axis=0)
    assert np.abs(m - post_mean).max() < 0.03
    # Euler-Maruyama with h = 1/||D||_F inflates the stationary spread (the reference shows the same bias:
    # SURVEY.md section 4 probe); the covariance must have the posterior's shape within that factor
    ratio = np.diag(c) / np.diag(post_cov)
    assert np.all(ratio > 0.8) and np.all(ratio < 2.5)
    assert abs(c[0, 1] / np.sqrt(c[0, 0] * c[1, 1]) - post_cov[0, 1] / np.sqrt(post_cov[0, 0] * post_cov[1, 1])) < 0.2


def test_flop_models():
    # SURVEY.md 8(d) quotes the dense-Gamma totals: cfg3 3.40e12, target 4.66e13
    assert abs(eo.algorithmic_flops(16384, 1024, 4096, gamma_dense=True) / 3.40e12 - 1) < 0.01
    assert abs(eo.algorithmic_flops(65536, 1024, 4096, gamma_dense=True) / 4.66e13 - 1) < 0.01
    assert abs(eo.algorithmic_flops(16384, 1024, 4096) / 2.852e12 - 1) < 0.01    # diagonal Gamma (bench workload)


def test_oracle_matches_golden_time_step_modes():
    """'constant', 'mix' and 'spectral' outputs of the real reference stored in tests/golden/timestep_cases.npz
    (tests/golden/make_golden_timestep.py): this pin travels to machines without the reference checkout."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "timestep_cases.npz"))
    for name in g["names"]:
        c = {key: g["%s/%s" % (name, key)] for key in ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi")}
        for rule in ("eks", "aldi"):
            for i, ms in enumerate(g["modes"]):
                mode, th = str(ms).split("|")
                th = [float(v) for v in th.split(",")] if th else None
                o = eo.step(rule, c["y"], c["U0"], c["G"], c["Gamma"], c["mu"], c["Sigma0"], c["ustar"], c["xi"],
                            time_step=mode, delta_t=0.03, t_last=th[-1] if th else None)
                Uk, (hk, t) = g["%s/%s/%d/Uk" % (name, rule, i)], g["%s/%s/%d/hk_t" % (name, rule, i)]
                assert np.abs(o["Uk"] - Uk).max() / np.abs(Uk).max() < TOL, (name, rule, ms)
                assert abs(o["hk"] - hk) <= 1e-12 * hk and abs(o["t"] - t) <= 1e-12 * t, (name, rule, ms)
