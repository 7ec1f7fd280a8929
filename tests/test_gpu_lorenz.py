"""Device Lorenz 63 / Lorenz 96 forward models (ces_b200/csrc/lorenz.cu) against the numpy restatement of the same
fixed-step scheme (oracle/lorenz_oracle.py, bit-level up to chaotic amplification of rounding) and, over a short horizon,
against golden trajectories of the real reference classes (tests/golden/lorenz_cases.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ces_b200 import calibrate, utils as cu  # noqa: E402
from oracle import lorenz_oracle as lo  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lorenz_cases.npz"))


def _ensemble(model, U, W0, t):
    Ud, Wd = torch.from_numpy(np.ascontiguousarray(U)).cuda(), torch.from_numpy(np.ascontiguousarray(W0)).cuda()
    G = torch.empty(model._n_stats(), U.shape[1], dtype=torch.float64, device="cuda")
    Wend = torch.empty_like(Wd)
    model.evaluate_ensemble_pde(None, Ud, Wd, t, G, Wend)
    return G.cpu().numpy(), Wend.cpu().numpy()


@pytest.mark.parametrize("log_params", [False, True])
def test_lorenz63_matches_oracle_and_reference(log_params):
    t = GOLD["l63_t"]
    model = (cu.lorenz63_log if log_params else cu.lorenz63)(l_window=1, freq=100)
    args = np.log(GOLD["l63_args"]) if log_params else GOLD["l63_args"]
    gold_ws = GOLD["l63log_ws" if log_params else "l63_ws"]
    gold_st = GOLD["l63log_stats" if log_params else "l63_stats"]
    G, Wend = _ensemble(model, args.T, GOLD["l63_w0"].T, t)
    for i in range(3):
        ws = lo.l63_solve(GOLD["l63_w0"][i], args[i], len(t), t[1] - t[0], model.substeps, log_params=log_params)
        # same scheme: rounding differences (FMA contraction) grow like exp(0.9 t), still ~1e-12 at T = 2
        assert np.abs(Wend[:, i] - ws[-1]).max() < 1e-9
        assert np.abs(G[:, i] - lo.l63_statistics(ws, 100)).max() < 1e-9 * np.abs(G[:, i]).max()
        # the reference's adaptive integrator over the same short horizon
        assert np.abs(Wend[:, i] - gold_ws[i][-1]).max() < 1e-3
        assert np.abs(G[:, i] - gold_st[i]).max() < 1e-4 * np.abs(gold_st[i]).max()
        # single-particle solve() returns the trajectory the statistics were taken from
        traj = model.solve(GOLD["l63_w0"][i], t, args=tuple(args[i]))
        assert traj.shape == (len(t), 3) and np.abs(traj - ws).max() < 1e-9
        assert np.abs(model.statistics(traj) - G[:, i]).max() < 1e-12 * np.abs(G[:, i]).max()


@pytest.mark.parametrize("cls,keys", [(cu.lorenz96, ("h", "F", "log_c", "b")), (cu.lorenz96Fc, ("F", "log_c")),
                                      (cu.lorenz96hcb, ("h", "log_c", "b"))])
def test_lorenz96_matches_oracle(cls, keys):
    ns, nf = 6, 4
    model = cls(n_slow=ns, n_fast=nf, l_window=1, freq=10, spinup=1) if cls is cu.lorenz96 else cls()
    model.n_slow, model.n_fast, model.n_state, model.l_window, model.freq, model.spinup = ns, nf, ns * (nf + 1), 1, 10, 1
    rng = np.random.default_rng(3)
    J = 5
    base = {"h": 1.0, "F": 10.0, "log_c": np.log(10.0), "b": 10.0}
    U = np.stack([np.array([base[k] for k in keys]) * (1.0 + 0.05 * rng.standard_normal(len(keys))) for _ in range(J)], axis=1)
    W0 = np.empty((model.n_state, J))
    for j in range(J):
        x = rng.random(ns) * 15 - 5
        W0[:, j] = np.concatenate([x, 0.05 * np.repeat(x, nf)])           # tame fast variables: slower error growth
    t = np.arange(0, 3.0 + 1e-9, 0.1)                                     # 31 samples: spin-up 11, two windows of 10
    G, Wend = _ensemble(model, U, W0, t)
    assert G.shape == (5 * ns, J)
    for j in range(J):
        traj = model.solve(W0[:, j], t, args=tuple(U[:, j]))
        # statistics and final state are those of the device's own trajectory (exact bookkeeping) ...
        assert np.abs(Wend[:, j] - traj[-1]).max() < 1e-12 * np.abs(traj).max()
        assert np.abs(G[:, j] - lo.l96_statistics(traj, ns, nf, 11, 10)).max() < 1e-11 * np.abs(G[:, j]).max()
        # ... and the trajectory is the oracle's while rounding has not yet been amplified (t <= 0.3)
        ws = lo.l96_solve(W0[:, j], ns, nf, dict(zip(keys, U[:, j])), 4, 0.1, model.substeps)
        assert np.abs(traj[:4] - ws).max() < 1e-9 * np.abs(ws).max()


def test_lorenz96_hom_outputs():
    hom = cu.lorenz96_hom()
    hom.n_slow, hom.n_fast, hom.n_state, hom.l_window, hom.freq, hom.spinup = 8, 4, 40, 1, 10, 1
    rng = np.random.default_rng(5)
    x = rng.random(8) * 15 - 5
    W0 = np.concatenate([x, 0.05 * np.repeat(x, 4)]).reshape(-1, 1)
    U = np.array([[1.0], [10.0], [np.log(10.0)], [10.0]])
    t = np.arange(0, 2.0 + 1e-9, 0.1)
    traj = hom.solve(W0[:, 0], t, args=tuple(U[:, 0]))
    for flag in (True, False):
        hom.hom = flag
        G, _ = _ensemble(hom, U, W0, t)
        assert G.shape == (5, 1)
        assert np.abs(G[:, 0] - hom.statistics(traj)).max() < 1e-11 * np.abs(G).max()


def test_run_with_device_lorenz63():
    """sampling.run on the device Lorenz 63 model (the lorenz63.ipynb scenario: parameters (r, b), statistics of a
    window, state carried over between iterations) agrees with the same loop driven through the reference's
    per-particle host protocol (G_pde_ens -> model.solve / model.statistics)."""
    model = cu.lorenz63(l_window=2, freq=50)
    model.substeps = 8
    t = np.arange(0, 4.0 + 1e-9, 0.02)              # 201 samples, two windows of 100
    wt = np.array([1.0, 2.0, 25.0])
    truth = model.solve(wt, t, args=(28.0, 8.0 / 3))
    y = model.statistics(truth)
    J, p = 24, 2
    Gamma = np.diag((0.05 * np.abs(y) + 0.1) ** 2)

    def sampler():
        s = calibrate.sampling(p, model.n_obs, J)
        s.ustar = np.array([[28.0], [8.0 / 3]])
        s.mu = np.array([[30.0], [3.0]])
        s.sigma = np.diag([25.0, 1.0])
        s.T = 3
        return s

    rng = np.random.default_rng(0)
    U0 = np.array([[30.0], [3.0]]) + np.array([[3.0], [0.5]]) * rng.standard_normal((p, J))
    xi = rng.standard_normal((p, J))
    a = sampler()
    a.run(y, U0, model, Gamma, None, t=t, wt=wt, xi=xi, t_tol=1e9)
    assert a.Uall.shape == (4, p, J) and a.Gall.shape == (4, model.n_obs + 3, J) and a.W0.shape == (3, J)
    assert np.isfinite(a.Ustar).all() and len(a.metrics["t"]) == 3

    class HostProtocol(object):            # the same model seen through the reference's per-particle interface only
        type, n_state, n_obs, model_name = 'pde', 3, model.n_obs, 'lorenz63'
        solve, statistics = staticmethod(model.solve), staticmethod(model.statistics)

    b = sampler()
    b.run(y, U0, HostProtocol(), Gamma, None, t=t, wt=wt, xi=xi, t_tol=1e9)
    # identical arithmetic per particle; chaotic amplification over 3 x 4 time units stays far below 1e-6
    assert np.abs(a.Gall[0] - b.Gall[0]).max() < 1e-9 * np.abs(b.Gall[0]).max()
    assert np.abs(a.Ustar - b.Ustar).max() < 1e-6 * np.abs(b.Ustar).max()


def test_lorenz63_example_script(monkeypatch):
    """examples/lorenz63_eks.py (the reference's lorenz63.ipynb scenario: data from a long trajectory, sampling.run with
    the state carried over, the notebook's direct G_pde_ens call on a parameter grid) runs end to end."""
    import importlib.util
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("lorenz63_example", os.path.join(root, "examples", "lorenz63_eks.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["lorenz63_eks.py", "--J", "32", "--T", "2", "--T-data", "60", "--T-eks", "10", "--grid", "5"])
    enki, Phi = mod.main()
    assert enki.Ustar.shape == (2, 32) and enki.Gstar.shape == (9, 32) and np.isfinite(enki.Ustar).all()
    assert enki.W0.shape == (3, 32) and len(enki.metrics["t"]) <= 2
    assert Phi.shape == (25,) and np.isfinite(Phi).all() and (Phi >= 0).all()


def test_g_pde_ens_batched_matches_per_particle():
    """enka.G_pde_ens on a device model (one launch) returns what the reference's per-particle loop over G_pde returns."""
    model = cu.lorenz63(l_window=1, freq=100)
    t = np.arange(0, 2.0 + 1e-9, 0.01)
    s = calibrate.sampling(2, 9, 3)
    theta = np.vstack([GOLD["l63_args"].T, GOLD["l63_w0"].T])
    got = s.G_pde_ens(theta, model, t)
    want = np.stack([s.G_pde(theta[:, j], model, t) for j in range(3)], axis=1)
    assert got.shape == (12, 3) and np.abs(got - want).max() < 1e-12 * np.abs(want).max()
