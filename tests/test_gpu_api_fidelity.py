"""API fidelity of the reference-facing classes beyond a single update:
resumed runs, the trace / online-save files, repeated single-particle model calls, ill-conditioned noise covariances."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ces_b200 import calibrate, utils as cutils  # noqa: E402
from oracle import eks_oracle as eo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def test_resumed_run_matches_the_real_reference():
    """A second run() on the same object continues Uall / Gall / metrics and the cumulative pseudo-time
    (ces/calibrate.py:307-315, 329-339), against outputs of the REAL reference (tests/golden/make_golden_resume.py)."""
    g = np.load(os.path.join(HERE, "golden", "resume_case.npz"))
    d, J = g["U0"].shape
    s = calibrate.sampling(d, g["A"].shape[0], J)
    s.ustar, s.mu, s.sigma, s.T = g["ustar"], g["mu"], g["Sigma0"], 3
    model = cutils.lineal(g["A"])
    np.random.seed(21)
    s.run(g["y"], g["U0"], model, g["Gamma"], None, t_tol=1e9)
    assert s.Uall.shape == g["first_Uall"].shape and np.allclose(s.metrics["t"], g["first_t"], rtol=1e-9, atol=0)
    s.T = 2
    s.run(g["y"], s.Ustar, model, g["Gamma"], None, t_tol=1e9)
    assert s.Uall.shape == g["Uall"].shape == (7, d, J) and s.Gall.shape == g["Gall"].shape
    scale = np.abs(g["Uall"]).max()
    assert np.abs(s.Uall - g["Uall"]).max() / scale < 1e-8          # 5 chained steps: rounding differences are amplified
    assert np.abs(s.Gall - g["Gall"]).max() / np.abs(g["Gall"]).max() < 1e-8
    for i, key in enumerate(("self-bias", "bias", "self-bias-data", "bias-data", "t")):
        assert len(s.metrics[key]) == 5 and np.allclose(s.metrics[key], g["metrics"][i], rtol=1e-7, atol=0), key
    assert np.all(np.diff(s.metrics["t"]) > 0)                        # the time keeps accumulating across the two calls
    assert np.abs(s.Ustar - g["Ustar"]).max() / scale < 1e-8 and np.allclose(s.Gstar, g["Gstar"], rtol=1e-6, atol=1e-8 * scale)


@pytest.mark.parametrize("fused", [True, False])
def test_banana_run_consumes_the_reference_random_stream(fused):
    """The reference's banana model draws two normals per evaluation even with its noise switched off (ces/utils.py:122),
    interleaved with the update noise of a seeded run: sampling.run on the device model reproduces the REAL reference's
    trace and leaves the generator in the same state (golden: tests/golden/make_golden_resume.py)."""
    g = np.load(os.path.join(HERE, "golden", "resume_case.npz"))
    s = calibrate.sampling(2, 2, 30)
    s.ustar, s.mu, s.sigma, s.T, s.fused_run = np.array([[0.4], [1.0]]), np.zeros((2, 1)), 9.0 * np.eye(2), 4, fused
    np.random.seed(8)
    s.run(np.array([0.4, 0.3]), g["banana_U0"], cutils.banana(), g["banana_Gamma"], None, t_tol=1e9)
    assert s.Uall.shape == g["banana_Uall"].shape
    assert np.abs(s.Uall - g["banana_Uall"]).max() / np.abs(g["banana_Uall"]).max() < 1e-8
    assert np.abs(s.Gall - g["banana_Gall"]).max() / np.abs(g["banana_Gall"]).max() < 1e-8
    assert np.allclose(s.metrics["t"], g["banana_t"], rtol=1e-8, atol=0)
    assert np.array_equal(np.random.normal(0, 1, 3), g["banana_next_normal"])


@pytest.mark.parametrize("J,d,k", [(60, 3, 7), (6000, 40, 24)])
def test_online_save_and_trace_files_hold_the_loop_states(tmp_path, J, d, k):
    """save_online=True writes ensemble_NNNN / Gensemble_NNNN (.npy) + metrics.pkl per iteration (ces/calibrate.py:371-385,
    170-197) through the asynchronous host trace (large ensembles: page-locked copies on a side stream; small ones: plain
    copies): iteration i's files hold the ensemble BEFORE update i, identical to the trace, and enka.load reads them back."""
    pr = eo.linear_gaussian_problem(d, k, J)
    s = calibrate.sampling(d, k, J)
    s.ustar, s.mu, s.sigma, s.T = pr["ustar"], pr["mu"], pr["Sigma0"], 4
    s.directory = str(tmp_path)
    model = cutils.lineal(pr["A"])
    model.l_window = 7
    np.random.seed(3)
    s.run(pr["y"], pr["U0"], model, pr["Gamma"], None, save_online=True, t_tol=1e9)
    where = os.path.join(str(tmp_path), "ensembles", "lineal-eks-007-%s" % str(J).zfill(4))
    assert s.Uall.shape == (5, d, J) and s.Gall.shape == (5, k, J)
    for i in range(4):
        Ui = np.load(os.path.join(where, "ensemble_%04d.npy" % i))
        Gi = np.load(os.path.join(where, "Gensemble_%04d.npy" % i))
        assert np.array_equal(Ui, s.Uall[i]) and np.array_equal(Gi, s.Gall[i])
    assert np.array_equal(s.Uall[0], pr["U0"])
    # the trace is the chain of oracle updates with the same noise stream
    np.random.seed(3)
    U, t = pr["U0"], None
    for i in range(4):
        o = eo.step("aldi", pr["y"], U, pr["A"] @ U, pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"],
                    np.random.normal(0, 1, [d, J]), t_last=t)
        U, t = o["Uk"], o["t"]
        assert np.abs(s.Uall[i + 1] - U).max() / np.abs(U).max() < 1e-9
    fresh = calibrate.sampling(d, k, J)
    assert fresh.load(path=os.path.join(str(tmp_path), "ensembles") + "/", eks_dir=os.path.basename(where) + "/", ix_ensemble=True)
    assert fresh.Uall.shape == (4, d, J) and np.array_equal(fresh.Uall, s.Uall[:4]) and fresh.metrics["t"] == s.metrics["t"]


def test_repeated_single_particle_calls_reuse_one_handle():
    """model(theta) in a loop (what ces/sample.py:121-196 does once per MCMC proposal) keeps ONE forward-only handle and
    its two device buffers on the model instead of creating an engine per call."""
    from oracle import forward_oracle as fo

    rng = np.random.default_rng(0)
    A = rng.standard_normal((9, 4))
    m = cutils.lineal(A, b=0.5)
    first = m(rng.standard_normal(4))
    eng = m._single_cache[1]
    for _ in range(50):
        th = rng.standard_normal(4)
        assert np.allclose(m(th), fo.lineal(A, th[:, None], 0.5)[:, 0], rtol=1e-13, atol=1e-13)
        assert m._single_cache[1] is eng
    assert first.shape == (9,)
    e = cutils.elliptic()
    x = e([-2.65, 104.5])
    assert np.allclose(x, [27.45194112300398, 79.70194112300398], rtol=1e-12)       # elliptic.ipynb:72
    import pickle

    m2 = pickle.loads(pickle.dumps(m))               # the cached handle does not travel (joblib, enka.parallel)
    assert np.allclose(m2(th), m(th), rtol=1e-14)


def _refined_solve(Gamma, R):
    """Gamma^-1 R to ~1e-16 relative even for cond(Gamma) = 1e8: Cholesky solve + iterative refinement with residuals in
    extended precision (np.longdouble, 80-bit on x86)."""
    L = np.linalg.cholesky(Gamma)
    solve = lambda B: np.linalg.solve(L.T, np.linalg.solve(L, B))
    X = solve(R)
    Gl, Rl = Gamma.astype(np.longdouble), R.astype(np.longdouble)
    for _ in range(4):
        res = (Rl - Gl @ X.astype(np.longdouble)).astype(np.float64)
        X = X + solve(res)
    return X


@pytest.mark.parametrize("cond", [1e2, 1e5, 1e8])
def test_ill_conditioned_dense_gamma_is_as_accurate_as_the_reference_solve(cond):
    """Dense Gamma goes through its Cholesky factor on the device (once) where the reference runs LAPACK's LU solve every
    step (ces/calibrate.py:461).  For cond(Gamma) >> 1 neither result can agree with the other to 1e-10 -- both carry a
    forward error ~ cond * eps -- so the check is against an extended-precision solve: the device path must be no less
    accurate than the reference's own arithmetic (numpy solve), and exact to 1e-10 when the conditioning allows it."""
    d, k, J = 12, 40, 90
    pr = eo.linear_gaussian_problem(d, k, J)
    rng = np.random.default_rng(4)
    Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    Gamma = (Q * np.geomspace(1.0, 1.0 / cond, k)) @ Q.T * 0.01
    Gamma = 0.5 * (Gamma + Gamma.T)
    E = pr["G"] - pr["G"].mean(axis=1, keepdims=True)
    R = pr["G"] - pr["y"][:, None]
    W = _refined_solve(Gamma, R)
    D = E.T @ W / J
    hk = 1.0 / (np.linalg.norm(D) + 1e-8)
    Ut = pr["U0"] - pr["U0"].mean(axis=1, keepdims=True)
    C = np.cov(pr["U0"]) + 1e-8 * np.eye(d)
    exact = (pr["U0"] - hk * (Ut @ D) - hk * (C @ np.linalg.solve(pr["Sigma0"], pr["U0"] - pr["mu"])) + hk * (d + 1.0) / J * Ut
             + np.sqrt(2 * hk) * (np.linalg.cholesky(C) @ pr["xi"]))
    ref = eo.step("aldi", pr["y"], pr["U0"], pr["G"], Gamma, pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
    s = calibrate.sampling(d, k, J)
    s.mu, s.sigma, s.ustar = pr["mu"], pr["Sigma0"], pr["ustar"]
    ours = s.eks_update_aldi(pr["y"], pr["U0"], pr["G"], Gamma, 0, xi=pr["xi"])
    scale = np.abs(exact).max()
    err_ours, err_ref = np.abs(ours - exact).max() / scale, np.abs(ref["Uk"] - exact).max() / scale
    assert err_ours <= max(10.0 * err_ref, 1e-10), (cond, err_ours, err_ref)
    if cond <= 1e2:
        assert err_ours < 1e-10


@pytest.mark.parametrize("kind,rule", [("lineal", "aldi"), ("lineal", "eks"), ("lineal", "aldi_constant"), ("lineal", "eki"),
                                       ("lineal_log", "aldi"), ("elliptic", "aldi"), ("banana", "eks")])
def test_fused_small_run_equals_the_iteration_by_iteration_loop(kind, rule):
    """ces_small_run (the whole run loop of a small problem in one launch) against the general loop of the same class:
    trace, forward outputs, metrics, stopping iteration and the state the global numpy generator is left in."""
    rs = np.random.RandomState(4)
    if kind in ("lineal", "lineal_log"):
        d, k, J = 2, 10, 100
        A = rs.normal(size=(k, d)) * (0.3 if kind == "lineal_log" else 1.0)
        model = cutils.lineal(A, b=0.25) if kind == "lineal" else cutils.lineal_log(A)
        ustar = np.array([[-1.0], [0.5]])
        y = model(ustar[:, 0]) + 0.05 * rs.normal(size=k)
        Gamma = 0.01 * np.eye(k) + (0.002 * np.ones((k, k)) if rule == "eks" else 0.0)      # dense Gamma for one rule
    else:
        d, k, J = 2, 2, 64
        model = cutils.elliptic() if kind == "elliptic" else cutils.banana()
        ustar = np.array([[-2.65], [104.5]]) if kind == "elliptic" else np.array([[0.5], [1.0]])
        y = np.asarray(model(ustar[:, 0]), dtype=float)
        Gamma = 0.01 * np.eye(2)
    U0 = ustar + rs.normal(size=(d, J))
    out = {}
    for fused in (True, False):
        s = calibrate.sampling(d, k, J)
        s.ustar, s.mu, s.sigma, s.T, s.fused_run = ustar, np.zeros((d, 1)) + ustar, 25.0 * np.eye(d), 25, fused
        np.random.seed(9)
        s.run(y, U0, model, Gamma, None, update=rule, t_tol=0.05)
        out[fused] = (s, np.random.get_state()[1].copy(), np.random.get_state()[2:])
    (a, sa, ta), (b, sb, tb) = out[True], out[False]
    n = len(b.metrics["t"])
    assert len(a.metrics["t"]) == n and 1 <= n <= 25 and a.Uall.shape == b.Uall.shape == (n + 1, d, J)
    scale = np.abs(b.Uall).max()
    assert np.abs(a.Uall - b.Uall).max() / scale < 1e-9 and np.abs(a.Gall - b.Gall).max() / max(np.abs(b.Gall).max(), 1e-300) < 1e-9
    for key in ("self-bias", "bias", "self-bias-data", "bias-data", "t"):
        assert np.allclose(a.metrics[key], b.metrics[key], rtol=1e-8, atol=0), key
    assert np.array_equal(a.Ustar, a.Uall[-1]) and np.array_equal(a.Gstar, a.Gall[-1]) and a.update_rule == b.update_rule
    assert np.array_equal(sa, sb) and ta == tb                      # the generator is where the reference would leave it
    # resume through the fused path: the time keeps accumulating, the trace grows
    a.T = 3
    a.run(y, a.Ustar, model, Gamma, None, update=rule, t_tol=1e9)
    assert a.Uall.shape[0] == n + 1 + 4 and len(a.metrics["t"]) == n + 3 and np.all(np.diff(a.metrics["t"]) > 0)
