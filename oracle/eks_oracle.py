"""numpy restatement of the EKS / ALDI particle update of agarbuno/ces.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- the checker for the CUDA
path and the timed CPU baseline, never the product.

Parity: PINNED.  tests/test_oracle.py checks every function here against the
real reference (``ces/calibrate.py`` run in this container through
oracle/reference_loader.py) and against tests/golden/step_*.npz generated from
it by tests/golden/make_golden.py.

Notation:  U (p,J) parameters, G (k,J) forward outputs,
particles on the contiguous axis; y (k,), Gamma (k,k) SPD, mu (p,1),
Sigma0 (p,p), ustar (p,1), xi (p,J) pre-drawn N(0,1).

Every function cites the reference lines it restates.  The restatement is
functional (no object state, noise passed in) where the reference is a method
with side effects on ``self.metrics`` and the global numpy RNG.
"""
import numpy as np

RULES = ("eks", "aldi", "aldi_constant", "eki")


def interaction(G, y, Gamma):
    """E, R and the J x J interaction matrix D = (1/J) E^T Gamma^{-1} R.

    ces/calibrate.py:427-429 (== :459-461 == :501-503)."""
    J = G.shape[1]
    E = G - G.mean(axis=1)[:, None]
    R = G - np.asarray(y)[:, None]
    W = np.linalg.solve(Gamma, R)
    D = (E.T @ W) / J
    return E, R, W, D


def metrics_cheap(U, ustar, E, R, Gamma):
    """The four per-step diagnostics, computed as column-wise quadratic forms.

    ces/calibrate.py:432-435 (== :464-467 == :506-509).  NB the two data-space
    metrics average the SQUARE of the Mahalanobis norms (``np.diag(..)**2``)."""
    Ut = U - U.mean(axis=1)[:, None]
    qe = np.einsum("ij,ij->j", E, np.linalg.solve(Gamma, E))
    qr = np.einsum("ij,ij->j", R, np.linalg.solve(Gamma, R))
    return {
        "self-bias": float((Ut ** 2).sum(axis=0).mean()),
        "bias": float(((U - np.asarray(ustar).reshape(U.shape[0], -1)) ** 2).sum(axis=0).mean()),
        "self-bias-data": float((qe ** 2).mean()),
        "bias-data": float((qr ** 2).mean()),
    }


def metrics_as_written(U, ustar, E, R, Gamma):
    """Same numbers with the reference's cost: two extra J x J products and two
    extra Gamma solves whose diagonals alone are used (ces/calibrate.py:434-435).
    Only the CPU-baseline timing uses this form."""
    Ut = U - U.mean(axis=1)[:, None]
    return {
        "self-bias": float((Ut ** 2).sum(axis=0).mean()),
        "bias": float(((U - np.asarray(ustar).reshape(U.shape[0], -1)) ** 2).sum(axis=0).mean()),
        "self-bias-data": float((np.diag(E.T @ np.linalg.solve(Gamma, E)) ** 2).mean()),
        "bias-data": float((np.diag(R.T @ np.linalg.solve(Gamma, R)) ** 2).mean()),
    }


def timestep(D, time_step=None, T=30, delta_t=None, t_last=None, spinup=4.0):
    """Step size rule.  ces/calibrate.py:247-260.

    ``t_last`` is the cumulative pseudo-time before this step (None: no step has
    been taken).  'adaptive' calls an undefined method in the reference
    (ces/calibrate.py:255) and is not restated."""
    frob = 1.0 / (np.linalg.norm(D) + 1e-8)
    const = delta_t if delta_t is not None else 1.0 / (T / 2)
    if time_step is None:
        return frob
    if time_step == "spectral":
        return 1.0 / np.linalg.eigvals(D).real.max()
    if time_step == "constant":
        return const
    if time_step == "mix":
        return frob if (t_last is None or t_last < spinup) else const
    raise ValueError("time_step=%r is not defined by the reference" % (time_step,))


def advance_time(hk, t_last):
    """Cumulative pseudo-time bookkeeping.  ces/calibrate.py:262-265."""
    return hk if t_last is None else hk + t_last


def _resolve_D(E, R, G, Gamma, hk, J):
    """D recomputed with Gamma -> hk*C^pp + Gamma.  ces/calibrate.py:439-441, :470-473."""
    Cpp = np.cov(G, bias=True).reshape(G.shape[0], G.shape[0])
    return (E.T @ np.linalg.solve(hk * Cpp + Gamma, R)) / J


def step(rule, y, U, G, Gamma, mu, Sigma0, ustar, xi, *, time_step=None, T=30, delta_t=None,
         t_last=None, spinup=4.0, switch=1.0, as_written=False, want_D=False):
    """One particle update.  Returns dict(Uk, hk, t, metrics[, D]).

    rule='eks'            ces/calibrate.py:418-449
    rule='aldi'           ces/calibrate.py:451-490   (the default, :304)
    rule='aldi_constant'  ces/calibrate.py:492-529
    rule='eki'            the deterministic part shared by all three,
                          U - h (U - ubar) D  (first two terms of :444 / :484);
                          the reference ships no EKI class.
    """
    U = np.asarray(U, dtype=float)
    G = np.asarray(G, dtype=float)
    p, J = U.shape
    mu = np.asarray(mu, dtype=float).reshape(p, -1)
    E, R, W, D = interaction(G, y, Gamma)
    mfun = metrics_as_written if as_written else metrics_cheap
    met = mfun(U, ustar, E, R, Gamma)
    if as_written and rule != "aldi_constant":
        np.linalg.cholesky(Gamma)  # the ignored Jnoise argument, :437 / :469
    ubar = U.mean(axis=1)[:, None]
    Ut = U - ubar

    if rule == "aldi_constant":
        C = np.cov(U).reshape(p, p) + 1e-8 * np.identity(p)             # :512
        alpha = (p + 1.0) / J                                            # :513
        drift = -(Ut @ D) - C @ np.linalg.solve(Sigma0, U - mu) + switch * alpha * Ut   # :515-517
        hk = 0.1 / np.max(np.abs(drift))                                 # :519
        t = advance_time(hk, t_last)                                     # :520-523
        Uk = U + hk * drift + np.sqrt(2 * hk) * (np.linalg.cholesky(C) @ xi)   # :525-527
    else:
        hk = timestep(D, time_step, T, delta_t, t_last, spinup)         # :437 / :469 (from the Gamma-only D)
        t = advance_time(hk, t_last)
        if rule == "eks":
            if time_step in ("adaptive", "constant"):                    # :439-441
                D = _resolve_D(E, R, G, Gamma, hk, J)
            C = np.cov(U, bias=True).reshape(p, p) + 1e-8 * np.identity(p)   # :424
            lhs = np.eye(p) + hk * np.linalg.solve(Sigma0.T, C.T).T      # :443
            rhs = U - hk * (Ut @ D) + hk * (C @ np.linalg.solve(Sigma0, mu))   # :444-445
            Uk = np.linalg.solve(lhs, rhs) + np.sqrt(2 * hk) * (np.linalg.cholesky(C) @ xi)   # :446-447
        elif rule == "aldi":
            if time_step in ("adaptive", "constant") or (time_step == "mix" and t > 1):   # :470-473
                D = _resolve_D(E, R, G, Gamma, hk, J)
            C = np.cov(U).reshape(p, p) + 1e-8 * np.identity(p)          # :476
            alpha = (p + 1.0) / J                                        # :481
            Uk = (U - hk * (Ut @ D) - hk * (C @ np.linalg.solve(Sigma0, U - mu))
                  + hk * alpha * Ut + np.sqrt(2 * hk) * (np.linalg.cholesky(C) @ xi))   # :484-488
        elif rule == "eki":
            Uk = U - hk * (Ut @ D)
        else:
            raise ValueError("unknown rule %r" % (rule,))
    out = {"Uk": Uk, "hk": float(hk), "t": float(t), "metrics": met}
    if want_D:
        out["D"] = D
    return out


def algorithmic_flops(J, d, k, gamma_dense=False):
    """W_step of SURVEY.md section 8(d): the flops the D-forming formulation needs,
    with no credit for the reference's redundant metric products."""
    w = 2.0 * k * J * J + 2.0 * d * J * J
    w += 2.0 * k * k * J if gamma_dense else 1.0 * k * J
    w += 6.0 * d * d * J + d ** 3 / 3.0
    return w


def reference_flops(J, d, k):
    """Flop model of the step as the reference writes it."""
    return 6.0 * k * J * J + 2.0 * d * J * J + 6.0 * k * k * J + 2.0 * k ** 3


def linear_gaussian_problem(d, k, J, seed=0, gamma=0.1, dense_gamma=False, noise_seed=1):
    """The synthetic problem of SURVEY.md section 8(d) / BASELINE.md section 3."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((k, d)) / np.sqrt(d)
    ustar = rng.standard_normal(d)
    if dense_gamma:
        Q = rng.standard_normal((k, k))
        Gamma = gamma ** 2 * (np.identity(k) + 0.5 * (Q @ Q.T) / k)
    else:
        Gamma = gamma ** 2 * np.identity(k)
    y = A @ ustar + gamma * rng.standard_normal(k)
    mu = np.zeros((d, 1))
    Sigma0 = 100.0 * np.identity(d)
    U0 = 10.0 * rng.standard_normal((d, J))
    G = A @ U0
    xi = np.random.RandomState(noise_seed).normal(0, 1, [d, J])
    return dict(A=A, ustar=ustar.reshape(d, 1), Gamma=Gamma, y=y, mu=mu, Sigma0=Sigma0, U0=U0, G=G, xi=xi)
