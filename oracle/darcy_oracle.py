"""scipy restatement of the Darcy forward model of agarbuno/ces.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  The reference evaluates this model through a MATLAB engine
(``import matlab.engine``, ces/darcy.py:4; ``eng.gaussrnd_coarse`` :92-95,138;
``eng.solve_gwf`` :97-98) that is not available in this environment (no MATLAB, no
Octave), and neither the repository nor its notebooks store a single Darcy
output.  This file restates the arithmetic of

    ces/darcy.py:20-38, 74-82, 84-98, 100-138      (Python side)
    utilities/mfiles/gaussrnd_coarse.m:6-23        (KL field by inverse 2-D DCT)
    utilities/mfiles/solve_gwf.m:4-39              (5-point finite-difference solve)

with the MATLAB builtins mapped as SURVEY.md section 8(c) documents:
``idct2`` -> ``scipy.fft.idctn(type=2, norm='ortho')``; ``interp2(..., 'spline')`` ->
separable not-a-knot cubic splines (``scipy.interpolate.CubicSpline``, extrapolating);
``spdiags``/``cell2mat`` assembly -> scipy.sparse; ``A\\F`` -> sparse direct solve;
``vec2mat`` + the final transpose cancel (the operator is index-symmetric).
"""
import numpy as np
import scipy.fft
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.interpolate import CubicSpline


def kl_eigs(N, alpha=2.0, tau=3.0):
    """sqrt-eigenvalues of the covariance operator on the N x N mode grid.
    gaussrnd_coarse.m:9,15 and ces/darcy.py:78-80 (identical formula; symmetric in K1, K2)."""
    k = np.arange(int(N))
    K1, K2 = np.meshgrid(k, k)
    return (tau ** (alpha - 1)) * (np.pi ** 2 * (K1 ** 2 + K2 ** 2) + tau ** 2) ** (-alpha / 2)


def set_rank(N, alpha=2.0, tau=3.0):
    """Order of the KL modes by decreasing eigenvalue, mode (0,0) last.  ces/darcy.py:74-82."""
    eigs = kl_eigs(N, alpha, tau)
    eigs[0, 0] = 1e-10
    return (-eigs).flatten().argsort()


def gaussrnd_coarse(xi, alpha, tau, N):
    """Log-permeability field at the N x N cell centres.  gaussrnd_coarse.m:6-23."""
    N = int(N)
    L = N * kl_eigs(N, alpha, tau) * np.asarray(xi, dtype=float).reshape(N, N)
    L[0, 0] = 0.0
    return scipy.fft.idctn(L, type=2, norm="ortho")


def _interp2_spline(xs, V, xq):
    """interp2(X, Y, V, Xq, Yq, 'spline') on identical tensor grids: not-a-knot cubic spline along
    each axis in turn (solve_gwf.m:13,37)."""
    T = CubicSpline(xs, V, axis=0, bc_type="not-a-knot", extrapolate=True)(xq)
    return CubicSpline(xs, T, axis=1, bc_type="not-a-knot", extrapolate=True)(xq)


def solve_gwf(theta):
    """Pressure at the cell centres for log-permeability theta (K x K).  solve_gwf.m:4-39."""
    K = theta.shape[0]
    centres = (np.arange(K) + 0.5) / K                      # :10
    nodes = np.arange(K) / (K - 1.0)                        # :11
    c = _interp2_spline(centres, np.exp(theta), nodes)      # :8,13  nodal coefficient
    n = K - 2
    idx = lambda i, j: (j - 1) * n + (i - 1)                # unknown of node (i, j), column blocks (:18-33)
    rows, cols, vals = [], [], []
    for j in range(1, K - 1):
        for i in range(1, K - 1):
            wn, ws = (c[i - 1, j] + c[i, j]) / 2, (c[i + 1, j] + c[i, j]) / 2
            ww, we = (c[i, j - 1] + c[i, j]) / 2, (c[i, j + 1] + c[i, j]) / 2
            rows.append(idx(i, j)); cols.append(idx(i, j)); vals.append(wn + ws + ww + we)     # :23-25
            if i > 1:
                rows.append(idx(i, j)); cols.append(idx(i - 1, j)); vals.append(-wn)           # :22
            if i < K - 2:
                rows.append(idx(i, j)); cols.append(idx(i + 1, j)); vals.append(-ws)           # :26
            if j > 1:
                rows.append(idx(i, j)); cols.append(idx(i, j - 1)); vals.append(-ww)           # :30-31
            if j < K - 2:
                rows.append(idx(i, j)); cols.append(idx(i, j + 1)); vals.append(-we)
    A = sp.csc_matrix((vals, (rows, cols)), shape=(n * n, n * n)) * (K - 1) ** 2               # :34
    x = spla.spsolve(A, np.ones(n * n))                                                        # :14-16,35 (F == 1)
    P = np.zeros((K, K))
    P[1:K - 1, 1:K - 1] = x.reshape(n, n).T          # x is column-block ordered: x[(j-1)n + (i-1)] = P(i, j)
    return _interp2_spline(nodes, P, centres)                                                  # :37


def nodal_coefficient(theta):
    K = theta.shape[0]
    return _interp2_spline((np.arange(K) + 0.5) / K, np.exp(theta), np.arange(K) / (K - 1.0))


class ModelTrunc(object):
    """ces/darcy.py:100-138 (``model_trunc``) without the MATLAB engine; ``p = None`` gives ``model``
    (:9-98, all N^2 coefficients)."""

    def __init__(self, alpha=2.0, tau=3.0, Nmesh=16, p=None):
        self.alpha, self.tau, self.Nmesh = alpha, tau, int(Nmesh)
        self.rank = set_rank(self.Nmesh, alpha, tau)
        self.p = self.Nmesh ** 2 if p is None else int(p)
        self.truncated = p is not None
        self.obs_index = None

    def eval_rf(self, xi):
        xi = np.asarray(xi, dtype=float)
        if self.truncated:                                   # :129-138
            full = np.zeros(self.Nmesh ** 2)
            full[self.rank[:self.p]] = xi
        else:                                                # :84-95
            full = xi
        return gaussrnd_coarse(full.reshape(self.Nmesh, -1), self.alpha, self.tau, self.Nmesh)

    def __call__(self, xi, full_solution=False):             # :20-38 / :111-122
        U = solve_gwf(self.eval_rf(xi)).flatten()
        return U if full_solution else U[self.obs_index]

    def set_initial(self, seed=1):                           # :66-72 / :124-127
        np.random.seed(seed)
        ustar = np.random.normal(0, 1, self.Nmesh ** 2)
        self.ustar = ustar[self.rank[:self.p]] if self.truncated else ustar
