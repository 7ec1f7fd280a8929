"""numpy restatement of the Lorenz 'pde'-type forward models of agarbuno/ces with the fixed-step RK4 scheme of
ces_b200/csrc/lorenz.cu.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Right-hand sides and statistics follow the reference line by line:
    lorenz63.model / __call__ / statistics        ces/utils.py:150-166, 181-194
    lorenz63_log.model                            ces/utils.py:207-221
    lorenz96.model / statistics                   ces/utils.py:289-308, 332-342
    lorenz96_hom.statistics                       ces/utils.py:354-368
The *integrator* is not the reference's: ces/utils.py:168-179 calls scipy's adaptive LSODA (odeint) and :316-330 an
adaptive RK45 with max_step; the device path uses `substeps` classical RK4 steps per output interval, and so does this
file, so that the device result can be compared step for step.  PARITY WITH THE REFERENCE INTEGRATORS IS PINNED ONLY OVER
SHORT HORIZONS (tests/golden/lorenz_cases.npz, made with the real ces.utils classes by tests/golden/make_golden_lorenz.py):
the systems are chaotic, so beyond a few Lyapunov times any two integrators decorrelate and only the statistics agree.
"""
import numpy as np


def l63_rhs(w, sigma, r, b):
    x, y, z = w
    return np.array([sigma * (y - x), r * x - y - x * z, x * y - b * z])          # ces/utils.py:163-166


def rk4(f, w, h):
    k1 = f(w)
    k2 = f(w + 0.5 * h * k1)
    k3 = f(w + 0.5 * h * k2)
    k4 = f(w + h * k3)
    return w + (h / 6.0) * ((k1 + k4) + 2.0 * (k2 + k3))


def integrate(f, w0, n_out, dt_out, substeps):
    """(n_out, n_state) samples at t_i = i dt_out."""
    w = np.asarray(w0, dtype=float).copy()
    out = np.empty((n_out, w.shape[0]))
    out[0] = w
    h = dt_out / substeps
    for i in range(1, n_out):
        for _ in range(substeps):
            w = rk4(f, w, h)
        out[i] = w
    return out


def l63_solve(w0, params, n_out, dt_out, substeps, log_params=False):
    r, b = (list(params) + [np.log(28.0) if log_params else 28.0, np.log(8.0 / 3) if log_params else 8.0 / 3])[:2] \
        if len(params) < 2 else params[:2]
    if len(params) == 1:
        b = np.log(8.0 / 3) if log_params else 8.0 / 3
    if log_params:
        r, b = np.exp(r), np.exp(b)                                              # :213-214
    return integrate(lambda w: l63_rhs(w, 10.0, r, b), w0, n_out, dt_out, substeps)


def l63_statistics(ws, window):
    """ces/utils.py:186-193: last adjacent window of t[1:]."""
    xs, ys, zs = ws[:, 0], ws[:, 1], ws[:, 2]
    m = np.asarray([xs, ys, zs, xs ** 2, ys ** 2, zs ** 2, xs * ys, xs * zs, ys * zs])
    return m[:, 1:].reshape(9, -1, window).mean(axis=2)[:, -1]


def l96_rhs(w, ns, nf, h, F, c, b):
    X, Y = w[:ns], w[ns:]
    n = ns * nf
    dX = np.empty(ns)
    for k in range(ns):                                                            # :298-301
        dX[k] = -X[k - 1] * (X[k - 2] - X[(k + 1) % ns]) - X[k] + F - (h * c) * np.mean(Y[k * nf:(k + 1) * nf])
    j = np.arange(n)
    dY = -c * b * Y[(j + 1) % n] * (Y[(j + 2) % n] - Y[j - 1]) - c * Y + ((h * c) / nf) * X[j // nf]     # :303-305
    return np.hstack((dX, dY))


def l96_solve(w0, ns, nf, params, n_out, dt_out, substeps):
    """params: dict with any of h, F, log_c, b (defaults of ces/utils.py:289)."""
    h, F, log_c, b = params.get("h", 1.0), params.get("F", 10.0), params.get("log_c", np.log(10.0)), params.get("b", 10.0)
    c = np.exp(log_c)
    return integrate(lambda w: l96_rhs(w, ns, nf, h, F, c, b), w0, n_out, dt_out, substeps)


def l96_statistics(ws, ns, nf, skip, window):
    """ces/utils.py:332-342 (skip = spinup * freq + 1)."""
    wsT = ws.T
    nstate = ns * (nf + 1)
    data = np.copy(wsT[:, skip:].reshape(nstate, -1, window))
    fast = data[ns:].reshape(ns, nf, -1, window)
    Phi = np.vstack([data[:ns].mean(axis=2), (data[:ns] ** 2).mean(axis=2), fast.mean(axis=1).mean(axis=2),
                     (data[ns:] ** 2).reshape(ns, nf, -1, window).mean(axis=1).mean(axis=2),
                     (data[:ns] * fast.mean(axis=1)).mean(axis=2)])
    return Phi[:, -1]
