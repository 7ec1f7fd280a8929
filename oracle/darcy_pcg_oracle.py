"""numpy / scipy.sparse restatement of the iterative solver inside ces_b200/csrc/darcy.cu.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference solves the 5-point system of utilities/mfiles/solve_gwf.m:16-35 with MATLAB's sparse direct solver; the
device runs conjugate gradients on the symmetrically Jacobi-scaled system A^ = S A S with the additive multilevel
preconditioner

    M^-1 = I + P1 diag(P1^T A^ P1)^-1 P1^T + Pc (Pc^T A^ Pc)^-1 Pc^T ,

P1 = piecewise constants on aggregates of 4 x 4 nodes, Pc = on aggregates of H x H nodes (H = 16 for Nmesh > 64, 8 for
Nmesh > 32, else 4; rows grouped from the first interior row, columns from the boundary column, exactly like the device's
thread tiles), stopping when r.M^-1 r <= tol^2 r0.M^-1 r0.  This file restates that algorithm so that the tests can (a)
check it against the direct solve of oracle/darcy_oracle.py on the CPU and (b) hold the device kernel to its iteration
counts: the preconditioner is part of the product's arithmetic and deserves an oracle of its own.
"""
import numpy as np
import scipy.sparse as sp

from . import darcy_oracle as do


def system(theta):
    """(A, n): the unscaled 5-point matrix of solve_gwf.m:16-34 without the (K-1)^2 factor (row-major interior nodes)
    for log-permeability theta (K x K); the right-hand side is h^2 = 1 / (K-1)^2 at every node."""
    K = theta.shape[0]
    c = do.nodal_coefficient(theta)
    n = K - 2
    I, J = np.meshgrid(np.arange(1, K - 1), np.arange(1, K - 1), indexing="ij")
    idx = lambda i, j: (i - 1) * n + (j - 1)
    wn, ws = (c[I - 1, J] + c[I, J]) / 2, (c[I + 1, J] + c[I, J]) / 2
    ww, we = (c[I, J - 1] + c[I, J]) / 2, (c[I, J + 1] + c[I, J]) / 2
    ii = idx(I, J)
    A = sp.coo_matrix(((wn + ws + ww + we).ravel(), (ii.ravel(), ii.ravel())), shape=(n * n, n * n)).tocsr()

    def off(mask, w, di, dj):
        return sp.coo_matrix((-w[mask], (ii[mask], idx(I[mask] + di, J[mask] + dj))), shape=(n * n, n * n)).tocsr()

    A = A + off(I > 1, wn, -1, 0) + off(I < K - 2, ws, 1, 0) + off(J > 1, ww, 0, -1) + off(J < K - 2, we, 0, 1)
    return A.tocsr(), n


def coarse_size(K):
    """Aggregate side of the coarse level (ces_b200/csrc/darcy.cu: coarse_geom on the template grid, the next multiple
    of 16 >= K): the smallest of 4, 8, 16 that leaves at most 8 x 8 aggregates."""
    Kt = 2 * (((K + 1) // 2 + 7) // 8 * 8)
    return 4 if Kt <= 32 else (8 if Kt <= 64 else 16)


def has_coarse_level(K):
    """Grids of at most 16 nodes a side run plain Jacobi scaling on the device (too few threads for the coarse solve)."""
    return 2 * (((K + 1) // 2 + 7) // 8 * 8) > 16


def aggregates(K, hr, hc):
    """Prolongator of piecewise constants: interior node (i, j) belongs to aggregate ((i-1) // hr, j // hc)."""
    n = K - 2
    I, J = np.meshgrid(np.arange(1, K - 1), np.arange(1, K - 1), indexing="ij")
    ar, ac = (I - 1) // hr, J // hc
    a = (ar * (ac.max() + 1) + ac).ravel()
    return sp.coo_matrix((np.ones(n * n), (np.arange(n * n), a)), shape=(n * n, a.max() + 1)).tocsr()


def solve(theta, tol=1e-13, max_iter=None, levels=("pair", "coarse")):
    """Nodal pressure (K x K, zero boundary) and the iteration count of the device's algorithm.
    ``levels``: which of the two levels above Jacobi scaling are active ("pair": 4 x 4 aggregates, "coarse": H x H)."""
    K = theta.shape[0]
    A, n = system(theta)
    s = 1.0 / np.sqrt(np.abs(A.diagonal()))
    Ah = (sp.diags(s) @ A @ sp.diags(s)).tocsr()
    b = s / float(K - 1) ** 2
    terms = []
    if not has_coarse_level(K) or A.diagonal().min() < 0.0:     # (a member with a negative diagonal runs plain scaling too)
        levels = ()
    if "pair" in levels and coarse_size(K) > 4:      # at Nmesh = 32 the coarse aggregates are these very blocks
        P1 = aggregates(K, 4, 4)
        terms.append((P1, 1.0 / np.asarray((P1.T @ Ah @ P1).diagonal()).ravel(), None))
    if "coarse" in levels:
        H = coarse_size(K)
        Pc = aggregates(K, H, H)
        terms.append((Pc, None, np.linalg.inv((Pc.T @ Ah @ Pc).toarray())))

    def precond(r):
        z = r.copy()
        for P, dinv, Binv in terms:
            rc = P.T @ r
            z = z + P @ (dinv * rc if Binv is None else Binv @ rc)
        return z

    max_iter = max_iter or 40 * K
    x = np.zeros_like(b)
    r = b.copy()
    z = precond(r)
    p = z.copy()
    rz = rz0 = r @ z
    it = 0
    for it in range(1, max_iter + 1):
        Ap = Ah @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = precond(r)
        rz_new = r @ z
        if rz_new <= tol * tol * rz0:
            break
        p = z + (rz_new / rz) * p
        rz = rz_new
    P = np.zeros((K, K))
    P[1:K - 1, 1:K - 1] = (s * x).reshape(n, n)
    return P, it
