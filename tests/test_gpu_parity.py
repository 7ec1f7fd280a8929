"""Parity of the CUDA update with the reference (FP64, relative tolerance 1e-10 per step as north_star
states): against the committed golden vectors of the real reference, against the numpy oracle on seeded
inputs, on the edge cases, through the reference-facing classes, and -- at BASELINE.json's cfg3 size --
through size-independent properties checked with torch fp64 on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ces_b200 import calibrate, utils as cutils  # noqa: E402
from ces_b200.engine import Engine  # noqa: E402
from oracle import eks_oracle as eo, forward_oracle as fo  # noqa: E402

TOL = 1e-10      # BASELINE.json north_star: FP64 relative tolerance per step
RULES = ("eks", "aldi", "aldi_constant")
METHOD = {"eks": "eks_update", "aldi": "eks_update_aldi", "aldi_constant": "eks_update_aldi_constant"}


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _sampler(d, k, J, mu, sigma, ustar, t_hist=None):
    s = calibrate.sampling(d, k, J)
    s.mu, s.sigma, s.ustar = mu, sigma, ustar
    if t_hist is not None and len(t_hist):
        s._ensure_metrics()
        s.metrics["t"] = list(t_hist)
    return s


def test_golden_vectors_of_the_real_reference(golden_steps):
    """sampling.eks_update* (numpy in, numpy out, noise from the seeded global numpy RNG exactly where
    the reference draws it) against outputs of the real reference stored in tests/golden."""
    g = golden_steps
    for name in g["names"]:
        c = {key: g["%s/%s" % (name, key)] for key in ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi", "t_hist")}
        d, J = c["U0"].shape
        for rule in RULES:
            s = _sampler(d, c["G"].shape[0], J, c["mu"], c["Sigma0"], c["ustar"], c["t_hist"])
            np.random.seed(1)                      # the golden xi is RandomState(1).normal(0, 1, [d, J])
            Uk = getattr(s, METHOD[rule])(c["y"], c["U0"], c["G"], c["Gamma"], 0)
            ref = g["%s/%s/Uk" % (name, rule)]
            m = g["%s/%s/metrics" % (name, rule)]
            assert _rel(Uk, ref) < TOL, (name, rule)
            got = [s.metrics[q][-1] for q in ("self-bias", "bias", "self-bias-data", "bias-data", "t")]
            assert np.allclose(got, m, rtol=TOL, atol=0), (name, rule, got, m)
            assert isinstance(Uk, np.ndarray) and Uk.shape == c["U0"].shape and Uk.flags["C_CONTIGUOUS"]


CASES = [(2, 10, 100), (64, 50, 1024), (40, 30, 17), (3, 5, 33), (1, 1, 2), (7, 1, 9), (1, 6, 250), (130, 257, 1000),
         (256, 512, 2048)]


@pytest.mark.parametrize("d,k,J", CASES)
@pytest.mark.parametrize("dense", [(False, False), (True, True), (True, False)])
def test_step_parity_against_oracle(d, k, J, dense):
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense[0])
    rng = np.random.default_rng(5)
    Sigma0, mu = pr["Sigma0"], pr["mu"]
    if dense[1]:
        S = rng.standard_normal((d, d))
        Sigma0, mu = 50 * np.eye(d) + S @ S.T, rng.standard_normal((d, 1))
    eng = Engine(d, k, J)
    try:
        eng.set_problem(pr["y"], pr["Gamma"], Sigma0, mu, pr["ustar"])
        for rule in RULES + ("eki",):
            o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], mu, Sigma0, pr["ustar"], pr["xi"])
            Uk, hk, met = eng.step_host(rule, pr["U0"], pr["G"], pr["xi"] if rule != "eki" else None)
            assert _rel(Uk, o["Uk"]) < TOL, rule
            assert abs(hk - o["hk"]) <= TOL * o["hk"], rule
            for key in met:
                assert abs(met[key] - o["metrics"][key]) <= TOL * abs(o["metrics"][key]), (rule, key)
    finally:
        eng.close()


def test_pipelined_host_step_panels_and_chunked_download():
    """ces_step_host on a shape that takes every pipelined branch: G uploaded in three row chunks with the first column
    panel of D accumulated over them (beta = 1, sum of squares from the stored values), several D panels (small
    d_panel_bytes), ragged J, and U_next downloaded in column chunks on the second stream.  Against the oracle, and
    against the device-resident step on the same inputs."""
    d, k, J = 24, 272, 4300
    pr = eo.linear_gaussian_problem(d, k, J)
    eng = Engine(d, k, J, d_panel_bytes=1536 * 8 * 4304)
    try:
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        for rule in RULES + ("eki",):
            xi = pr["xi"] if rule != "eki" else None
            o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
            Uk, hk, met = eng.step_host(rule, pr["U0"], pr["G"], xi)
            assert _rel(Uk, o["Uk"]) < TOL and abs(hk - o["hk"]) <= TOL * o["hk"], rule
            for key in met:
                assert abs(met[key] - o["metrics"][key]) <= TOL * abs(o["metrics"][key]), (rule, key)
            Ud, hd, _ = eng.step(rule, dev(pr["U0"]), dev(pr["G"]), dev(xi) if xi is not None else None)
            assert _rel(Uk, Ud.cpu().numpy()) < 1e-12 and abs(hk - hd) <= 1e-13 * hd, rule
        # pageable and page-locked host arrays give the same result
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        U2, h2, _ = eng.step_host("aldi", pin(pr["U0"]), pin(pr["G"]), pin(pr["xi"]))
        U1, h1, _ = eng.step_host("aldi", pr["U0"], pr["G"], pr["xi"])
        assert np.array_equal(U1, U2) and h1 == h2
    finally:
        eng.close()


@pytest.mark.parametrize("d,k,J,dense", [(6, 9, 40, True), (20, 33, 150, False), (64, 130, 600, True)])
def test_non_default_time_steps_match_oracle(d, k, J, dense):
    """time_step='constant' / 'mix' with the hk C^pp + Gamma re-solve of D (ces/calibrate.py:439-441, 470-473)
    through the reference-facing methods, against the oracle (itself pinned to the real reference for these modes)."""
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense)
    for rule in ("aldi", "eks"):
        for ts, t_hist in (("constant", []), ("constant", [0.4, 0.9]), ("mix", []), ("mix", [0.5, 1.7]), ("mix", [2.0, 5.0])):
            s = _sampler(d, k, J, pr["mu"], pr["Sigma0"], pr["ustar"], t_hist)
            s._ensure_metrics()
            np.random.seed(1)
            Uk = getattr(s, METHOD[rule])(pr["y"], pr["U0"], pr["G"], pr["Gamma"], 0, time_step=ts, delta_t=0.03)
            o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"],
                        time_step=ts, delta_t=0.03, T=s.T, t_last=t_hist[-1] if t_hist else None)
            assert _rel(Uk, o["Uk"]) < TOL, (rule, ts, t_hist)
            assert abs(s.metrics["t"][-1] - o["t"]) < TOL * o["t"], (rule, ts, t_hist)


@pytest.mark.parametrize("d,k,J,dense", [(2, 10, 100, False), (6, 9, 40, True), (20, 33, 17, False), (64, 130, 600, True),
                                         (64, 700, 2000, False), (16, 600, 900, True)])
def test_spectral_time_step_matches_oracle(d, k, J, dense):
    """time_step='spectral' (ces/calibrate.py:249-251): hk = 1 / eigvals(D).real.max().  The device never forms an
    eigen-decomposition of the J x J matrix: lambda_max(D) = lambda_max(Gamma^-1 C^pp), found by Lanczos (csrc/eig.cu);
    the oracle calls numpy's non-symmetric eigvals on D like the reference."""
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense)
    for rule in ("aldi", "eks"):
        s = _sampler(d, k, J, pr["mu"], pr["Sigma0"], pr["ustar"])
        s._ensure_metrics()
        Uk = getattr(s, METHOD[rule])(pr["y"], pr["U0"], pr["G"], pr["Gamma"], 0, time_step="spectral", xi=pr["xi"])
        o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"],
                    time_step="spectral")
        assert abs(s.metrics["t"][-1] - o["hk"]) < TOL * o["hk"], rule
        assert abs(s.radspec[-1] - 1.0 / o["hk"]) < TOL / o["hk"], rule
        assert _rel(Uk, o["Uk"]) < TOL, rule


@pytest.mark.parametrize("d,k,J,dense", [(2, 10, 100, False), (40, 30, 17, True), (64, 50, 1024, False), (130, 257, 1000, True),
                                         (1024, 10, 300, False)])
def test_factored_formulation_matches_oracle(d, k, J, dense):
    """formulation='factored' (D never formed) is the same update to rounding."""
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense)
    for rule in RULES + ("eki",):
        s = _sampler(d, k, J, pr["mu"], pr["Sigma0"], pr["ustar"])
        o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
        name = METHOD.get(rule, "eki_update")
        Uk = getattr(s, name)(pr["y"], pr["U0"], pr["G"], pr["Gamma"], 0, xi=pr["xi"], formulation="factored")
        assert _rel(Uk, o["Uk"]) < TOL, rule
        assert abs(s.metrics["t"][-1] - o["hk"]) < TOL * o["hk"], rule
        for key in ("self-bias", "bias", "self-bias-data", "bias-data"):
            assert abs(s.metrics[key][-1] - o["metrics"][key]) <= TOL * abs(o["metrics"][key])


def test_device_resident_step_with_strided_tensors():
    """Engine.step on CUDA tensors that are column slices of wider buffers (leading dimension != J, odd
    offset so the noise operand needs the internal re-pack)."""
    d, k, J = 20, 12, 101
    pr = eo.linear_gaussian_problem(d, k, J)
    eng = Engine(d, k, J)
    try:
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        wide = lambda a: torch.from_numpy(np.concatenate([np.zeros((a.shape[0], 3)), a, np.zeros((a.shape[0], 5))], axis=1)).cuda()[:, 3:3 + J]
        U, G, xi = wide(pr["U0"]), wide(pr["G"]), wide(pr["xi"])
        out, hk, met = eng.step("aldi", U, G, xi)
        o = eo.step("aldi", pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
        assert _rel(out.cpu().numpy(), o["Uk"]) < TOL
        assert torch.equal(U.cpu(), torch.from_numpy(pr["U0"]))      # inputs are never mutated
    finally:
        eng.close()


def test_d_panel_streaming_is_exact():
    """A tiny D workspace forces several column panels; the result must not change."""
    d, k, J = 33, 21, 700
    pr = eo.linear_gaussian_problem(d, k, J)
    o = eo.step("aldi", pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
    eng = Engine(d, k, J, d_panel_bytes=128 * 8 * 704)       # 128-column panels -> 6 panels
    try:
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        Uk, hk, _ = eng.step_host("aldi", pr["U0"], pr["G"], pr["xi"])
        assert _rel(Uk, o["Uk"]) < TOL and abs(hk - o["hk"]) < TOL * hk
    finally:
        eng.close()


def test_errors_follow_numpy_conventions():
    s = calibrate.sampling(3, 4, 10)
    y, U, G, Gam = np.zeros(4), np.random.default_rng(0).standard_normal((3, 10)), np.ones((4, 10)), np.eye(4)
    with pytest.raises(AttributeError):                # mu / sigma / ustar unset (ces/calibrate.py:433, 443)
        s.eks_update_aldi(y, U, G, Gam, 0)
    s.mu, s.ustar = np.zeros((3, 1)), np.zeros((3, 1))
    s.sigma = -np.eye(3)
    with pytest.raises(np.linalg.LinAlgError):
        s.eks_update_aldi(y, U, G, Gam, 0)
    s.sigma = np.eye(3)
    bad = np.array([[1.0, 2.0, 0, 0], [2.0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]])
    with pytest.raises(np.linalg.LinAlgError):
        s.eks_update_aldi(y, U, G, bad, 0)


def test_forward_maps_match_the_reference_models(golden_forward):
    g = golden_forward
    e = calibrate.enka(4, 10, 37)
    assert _rel(e.G_ens(g["lineal/U"], cutils.lineal(g["lineal/A"], b=float(g["lineal/b"]))), g["lineal/G"]) < 1e-13
    assert _rel(e.G_ens(g["lineal/U"], cutils.lineal_log(g["lineal/A"])), g["lineal_log/G"]) < 1e-13
    e2 = calibrate.enka(2, 2, 53)
    assert _rel(e2.G_ens(g["map2/U"], cutils.elliptic()), g["elliptic/G"]) < 1e-12
    assert _rel(e2.G_ens(g["map2/U"], cutils.banana(a=1.3, b=0.4)), g["banana/G"]) < 1e-13
    # single-particle call, as user scripts do to make y_obs (examples/notebooks/elliptic.ipynb:72)
    y = cutils.elliptic()(g["elliptic/ustar_notebook"])
    assert np.allclose(y, g["elliptic/y_obs_notebook"], rtol=1e-12)
    single = cutils.lineal(g["lineal/A"], b=float(g["lineal/b"]))(g["lineal/U"][:, 3])
    assert _rel(single, g["lineal/G"][:, 3]) < 1e-13


@pytest.mark.parametrize("rule", ["aldi", "eks", "aldi_constant"])
def test_run_loop_matches_oracle_loop(rule):
    """sampling.run end to end (device forward model, ensemble resident in HBM, metrics, t_tol stop, trace)
    against the oracle stepped in a loop on the same seeded numpy random stream."""
    np.random.seed(1)
    A = np.ones((10, 2))
    A[:, 1] = 2 * np.random.normal(0, 1, 10)
    ustar = np.array([[-1.0], [2.0]])
    Gamma = 0.1 * np.eye(10)
    y = A @ ustar[:, 0] + np.sqrt(0.1) * np.random.normal(0, 1, 10)
    J, T = 100, 25
    U0 = 3.0 * np.random.normal(0, 1, [2, J])
    s = calibrate.sampling(2, 10, J)
    s.ustar, s.mu, s.sigma, s.T = ustar, np.zeros((2, 1)), 100.0 * np.eye(2), T
    np.random.seed(7)
    s.run(y, U0, cutils.lineal(A), Gamma, np.linalg.cholesky(Gamma), update=rule, t_tol=0.35)
    np.random.seed(7)
    U, t, n = U0, None, 0
    for _ in range(T):
        xi = np.random.normal(0, 1, [2, J])
        o = eo.step(rule, y, U, fo.lineal(A, U), Gamma, s.mu, s.sigma, ustar, xi, t_last=t)
        U, t, n = o["Uk"], o["t"], n + 1
        if t > 0.35:
            break
    assert len(s.metrics["t"]) == n
    assert abs(s.metrics["t"][-1] - t) < 1e-9 * t
    assert _rel(s.Ustar, U) < 1e-8                      # n chained steps, each within 1e-10
    assert s.Uall.shape == (n + 1, 2, J) and s.Gall.shape == (n + 1, 10, J)
    assert np.array_equal(s.Uall[0], U0) and _rel(s.Gstar, fo.lineal(A, s.Ustar)) < 1e-12
    assert s.update_rule == {"aldi": "eks_update_linear", "eks": "eks_update", "aldi_constant": "eks_update_aldi"}[rule]


def test_run_with_a_user_callable_keeps_the_reference_protocol():
    class Cubic(object):
        type = "map"
        model_name = "cubic"

        def __call__(self, theta):
            return np.array([theta[0] ** 3 + theta[1], theta[0] - theta[1], theta[1] ** 2])

    J = 40
    rng = np.random.default_rng(2)
    U0 = rng.standard_normal((2, J))
    s = calibrate.sampling(2, 3, J)
    s.ustar, s.mu, s.sigma, s.T = np.array([[0.5], [0.2]]), np.zeros((2, 1)), 4.0 * np.eye(2), 3
    y, Gamma = np.array([0.3, 0.3, 0.04]), 0.01 * np.eye(3)
    np.random.seed(11)
    s.run(y, U0, Cubic(), Gamma, None, t_tol=1e9)
    np.random.seed(11)
    U, t = U0, None
    for _ in range(3):
        G = np.stack([Cubic()(c) for c in U.T], axis=1)
        o = eo.step("aldi", y, U, G, Gamma, s.mu, s.sigma, s.ustar, np.random.normal(0, 1, [2, J]), t_last=t)
        U, t = o["Uk"], o["t"]
    assert _rel(s.Ustar, U) < 1e-9 and len(s.Uall) == 4


# ---------------------------------------------------------------------------------------------------
# BASELINE.json cfg3 size (d=1024, k=4096, J=16384): the numpy oracle needs minutes and ~10 GB here, so
# the full size is checked through properties, with torch fp64 (cuBLAS) on the device as the reference.
@pytest.fixture(scope="module")
def cfg3():
    d, k, J = 1024, 4096, 16384
    gen = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=gen)
    A = rn(k, d) / d ** 0.5
    ustar = rn(d)
    y = A @ ustar + 0.1 * rn(k)
    U = 10.0 * rn(d, J)
    G = A @ U
    xi = rn(d, J)
    eng = Engine(d, k, J)
    eng.set_problem(y.cpu().numpy(), 0.01 * np.eye(k), 100.0 * np.eye(d), np.zeros(d), ustar.cpu().numpy())
    out, hk, met = eng.step("aldi", U, G, xi)
    yield dict(d=d, k=k, J=J, y=y, U=U, G=G, xi=xi, ustar=ustar, eng=eng, out=out, hk=hk, met=met)
    eng.close()


def test_full_size_step_size_via_gram_identity(cfg3):
    """||D||_F^2 = sum((E E^T) o (W W^T)) / J^2 -- two k x k Gram matrices instead of the J x J matrix."""
    c = cfg3
    E = c["G"] - c["G"].mean(dim=1, keepdim=True)
    W = (c["G"] - c["y"][:, None]) / 0.01
    frob2 = float(((E @ E.t()) * (W @ W.t())).sum()) / c["J"] ** 2
    hk = 1.0 / (frob2 ** 0.5 + 1e-8)
    assert abs(c["hk"] - hk) < 1e-10 * hk
    qr = ((c["G"] - c["y"][:, None]) * W).sum(dim=0)
    assert abs(c["met"]["bias-data"] - float((qr ** 2).mean())) < 1e-10 * c["met"]["bias-data"]
    assert abs(c["met"]["bias"] - float(((c["U"] - c["ustar"][:, None]) ** 2).sum(dim=0).mean())) < 1e-10 * c["met"]["bias"]


def test_full_size_column_probe(cfg3):
    """256 random particles of U_{n+1} recomputed from the definition (ces/calibrate.py:459-488)."""
    c = cfg3
    d, J, h = c["d"], c["J"], c["hk"]
    cols = torch.randperm(J, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))[:256]
    U, G = c["U"], c["G"]
    E = G - G.mean(dim=1, keepdim=True)
    Ut = U - U.mean(dim=1, keepdim=True)
    Wc = (G[:, cols] - c["y"][:, None]) / 0.01
    Dc = (E.t() @ Wc) / J
    C = (Ut @ Ut.t()) / (J - 1) + 1e-8 * torch.eye(d, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(C)
    ref = (U[:, cols] - h * (Ut @ Dc) - h * (C @ (U[:, cols] / 100.0)) + h * (d + 1.0) / J * Ut[:, cols]
           + (2 * h) ** 0.5 * (L @ c["xi"][:, cols]))
    got = c["out"][:, cols]
    assert float((got - ref).abs().max() / ref.abs().max()) < TOL


def test_full_size_permutation_equivariance(cfg3):
    """Relabelling the particles relabels the update: U+(P) = U+ P (sums over particles change order only)."""
    c = cfg3
    perm = torch.randperm(c["J"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    out2, hk2, _ = c["eng"].step("aldi", c["U"][:, perm].contiguous(), c["G"][:, perm].contiguous(),
                                 c["xi"][:, perm].contiguous())
    assert abs(hk2 - c["hk"]) < 1e-11 * c["hk"]
    assert float((out2 - c["out"][:, perm]).abs().max() / c["out"].abs().max()) < TOL


class _DampedOscillator(object):
    """A 'pde'-type model in the reference's protocol (ces/utils.py:124-194 is the Lorenz-63 instance of it):
    solve(w0, t, args) integrates an ODE, statistics(ws) reduces the trajectory, n_state is the carried state."""
    type = "pde"
    model_name = "oscillator"
    n_state = 2
    n_obs = 3

    def solve(self, w0, t, args=()):
        k, c = args
        ws = np.empty((len(t), 2))
        ws[0] = w0
        for n in range(1, len(t)):
            h = t[n] - t[n - 1]
            x, v = ws[n - 1]
            ws[n] = [x + h * v, v + h * (-np.exp(k) * x - np.exp(c) * v + 1.0)]
        return ws

    def statistics(self, ws):
        return np.array([ws[:, 0].mean(), ws[:, 1].mean(), (ws[:, 0] ** 2).mean()])


@pytest.mark.parametrize("use_pool", [False, True])
def test_run_with_a_pde_type_model(use_pool):
    """'pde' branch of sampling.run (ces/calibrate.py:317-327, 342-350, 390-396): state carry-over W0, the ws pool
    with its np.random.randint draws interleaved with the noise draws, Gall holding statistics + final state."""
    model, J, T, p = _DampedOscillator(), 30, 4, 2
    rng = np.random.default_rng(0)
    U0 = 0.3 * rng.standard_normal((p, J))
    y, Gamma = np.array([0.4, 0.0, 0.3]), 0.01 * np.eye(3)
    t = np.linspace(0.0, 2.0, 41)
    wt = np.array([0.5, 0.0])
    pool = rng.standard_normal((17, 2)) if use_pool else None
    s = calibrate.sampling(p, 3, J)
    s.ustar, s.mu, s.sigma, s.T = np.zeros((p, 1)), np.zeros((p, 1)), 4.0 * np.eye(p), T
    np.random.seed(21)
    s.run(y, U0, model, Gamma, None, wt=wt, t=t, ws=pool, t_tol=1e9)
    # the reference protocol, stepped with the oracle
    np.random.seed(21)
    e = calibrate.enka(p, 3, J)
    if pool is not None:
        W0 = pool[np.random.randint(pool.shape[0], size=J)].T
    else:
        W0 = np.tile(wt, J).reshape(J, 2).T
    U, tt = U0, None
    for _ in range(T):
        G = e.G_pde_ens(np.vstack([U, W0]), model, t)
        W0 = pool[np.random.randint(pool.shape[0], size=J)].T if pool is not None else np.copy(G[3:, :])
        o = eo.step("aldi", y, U, G[:3], Gamma, s.mu, s.sigma, s.ustar, np.random.normal(0, 1, [p, J]), t_last=tt)
        U, tt = o["Uk"], o["t"]
    G = e.G_pde_ens(np.vstack([U, W0]), model, t)
    assert _rel(s.Ustar, U) < 1e-8
    assert s.Gall.shape == (T + 1, 3 + 2, J) and _rel(s.Gall[-1], G) < 1e-8
    assert _rel(s.Gstar, G[:3]) < 1e-8
    if pool is not None:
        assert len(s.Wall) == T + 1


@pytest.mark.parametrize("d,k,J", [(2, 10, 100), (3, 5, 33), (8, 16, 512), (1, 1, 2)])
def test_general_path_on_small_shapes(d, k, J, monkeypatch):
    """Small problems normally take the single-kernel path (small.cu); the general multi-kernel path must give the
    same answers on them (CES_NO_SMALL_PATH=1 is read when the handle is created)."""
    monkeypatch.setenv("CES_NO_SMALL_PATH", "1")
    pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=True)
    eng = Engine(d, k, J)
    monkeypatch.delenv("CES_NO_SMALL_PATH")
    eng2 = Engine(d, k, J)
    try:
        for e in (eng, eng2):
            e.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        for rule in RULES + ("eki",):
            o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
            n0 = eng.launch_count()
            Ua, ha, _ = eng.step_host(rule, pr["U0"], pr["G"], pr["xi"])
            n1 = eng.launch_count()
            Ub, hb, mb = eng2.step_host(rule, pr["U0"], pr["G"], pr["xi"])
            n2 = eng.launch_count()
            assert n2 - n1 == 1 and n1 - n0 > 5          # one kernel vs the general sequence
            assert _rel(Ua, o["Uk"]) < TOL and _rel(Ub, o["Uk"]) < TOL, rule
            assert abs(ha - o["hk"]) < TOL * o["hk"] and abs(hb - o["hk"]) < TOL * o["hk"]
            for key in mb:
                assert abs(mb[key] - o["metrics"][key]) <= TOL * abs(o["metrics"][key]), (rule, key)
    finally:
        eng.close()
        eng2.close()


def test_split_contraction_of_the_drift_product():
    """d = 1024, J = 8192 gives 8 x 64 = 512 tiles for V = U~ D (3.46 waves of 148 SMs): the library splits the
    contraction; the result must not change beyond rounding."""
    d, k, J = 1024, 24, 8192
    pr = eo.linear_gaussian_problem(d, k, J)
    o = eo.step("aldi", pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
    eng = Engine(d, k, J)
    try:
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        Uk, hk, met = eng.step_host("aldi", pr["U0"], pr["G"], pr["xi"])
        assert _rel(Uk, o["Uk"]) < TOL and abs(hk - o["hk"]) < TOL * hk
    finally:
        eng.close()


def test_timestep_method_with_an_explicit_interaction_matrix():
    """sampling.timestep_method(D, ...) for callers that own D (ces/calibrate.py:243-267): default rule, the time
    bookkeeping, 'constant' and 'mix'."""
    rng = np.random.default_rng(3)
    D = rng.standard_normal((37, 37))
    s = calibrate.sampling(2, 3, 37)
    s.T = 30
    s._ensure_metrics()
    h1 = s.timestep_method(D, None, None, None, None)
    assert abs(h1 - eo.timestep(D)) < 1e-14 * h1 and s.metrics["t"] == [h1]
    h2 = s.timestep_method(D, None, None, None, None, time_step="constant")
    assert h2 == 1. / 15 and abs(s.metrics["t"][-1] - (h1 + h2)) < 1e-15
    h3 = s.timestep_method(D, None, None, None, None, time_step="mix", spinup=0.01, delta_t=0.25)
    assert h3 == 0.25
    # 'spectral' on a caller-owned, non-symmetric D (ces/calibrate.py:249-251): radspec = eigvals(D).real.max()
    t_before = s.metrics["t"][-1]
    h4 = s.timestep_method(D, None, None, None, None, time_step="spectral")
    lam = np.linalg.eigvals(D).real.max()
    assert abs(h4 - 1.0 / lam) < 1e-10 * abs(h4) and abs(s.radspec[-1] - lam) < 1e-10 * abs(lam)
    assert abs(s.metrics["t"][-1] - (t_before + h4)) < 1e-15
    with pytest.raises(NotImplementedError):
        s.timestep_method(D, None, None, None, None, time_step="adaptive")


def test_target_size_properties():
    """BASELINE.json's target shape (d=1024, k=4096, J=65536; D is formed in 4 column panels): step size through
    the Gram identity and a 128-particle probe of U_{n+1} recomputed from the definition with torch fp64 (cuBLAS)."""
    d, k, J = 1024, 4096, 65536
    gen = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=gen)
    A = rn(k, d) / d ** 0.5
    ustar = rn(d)
    y = A @ ustar + 0.1 * rn(k)
    U = 10.0 * rn(d, J)
    G = A @ U
    xi = rn(d, J)
    eng = Engine(d, k, J)
    try:
        eng.set_problem(y.cpu().numpy(), 0.01 * np.eye(k), 100.0 * np.eye(d), np.zeros(d), ustar.cpu().numpy())
        out, hk, met = eng.step("aldi", U, G, xi)
        E = G - G.mean(dim=1, keepdim=True)
        W = (G - y[:, None]) / 0.01
        frob2 = float(((E @ E.t()) * (W @ W.t())).sum()) / J ** 2
        h_ref = 1.0 / (frob2 ** 0.5 + 1e-8)
        assert abs(hk - h_ref) < 1e-10 * h_ref
        cols = torch.randperm(J, device="cuda", generator=gen)[:128]
        Ut = U - U.mean(dim=1, keepdim=True)
        Dc = (E.t() @ W[:, cols]) / J
        C = (Ut @ Ut.t()) / (J - 1) + 1e-8 * torch.eye(d, dtype=torch.float64, device="cuda")
        L = torch.linalg.cholesky(C)
        ref = (U[:, cols] - hk * (Ut @ Dc) - hk * (C @ (U[:, cols] / 100.0)) + hk * (d + 1.0) / J * Ut[:, cols]
               + (2 * hk) ** 0.5 * (L @ xi[:, cols]))
        assert float((out[:, cols] - ref).abs().max() / ref.abs().max()) < TOL
    finally:
        eng.close()


def test_golden_time_step_modes_of_the_real_reference():
    """time_step = 'constant' / 'mix' / 'spectral' through the reference-facing methods against outputs of the REAL
    reference (tests/golden/timestep_cases.npz), not just against the oracle."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "timestep_cases.npz"))
    for name in g["names"]:
        c = {key: g["%s/%s" % (name, key)] for key in ("y", "U0", "G", "Gamma", "mu", "Sigma0", "ustar", "xi")}
        d, J = c["U0"].shape
        for rule in ("eks", "aldi"):
            for i, ms in enumerate(g["modes"]):
                mode, th = str(ms).split("|")
                th = [float(v) for v in th.split(",")] if th else []
                s = _sampler(d, c["G"].shape[0], J, c["mu"], c["Sigma0"], c["ustar"], th)
                s._ensure_metrics()
                Uk = getattr(s, METHOD[rule])(c["y"], c["U0"], c["G"], c["Gamma"], 0, time_step=mode, delta_t=0.03, xi=c["xi"])
                ref, (hk, t) = g["%s/%s/%d/Uk" % (name, rule, i)], g["%s/%s/%d/hk_t" % (name, rule, i)]
                assert _rel(Uk, ref) < TOL, (name, rule, ms)
                assert abs(s.metrics["t"][-1] - t) < TOL * t, (name, rule, ms)
