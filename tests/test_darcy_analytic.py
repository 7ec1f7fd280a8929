"""Analytic (manufactured-solution) checks of the Darcy discretisation -- what can be pinned WITHOUT the MATLAB original.

The reference's Darcy model runs through a MATLAB engine (ces/darcy.py:4, 92-98) that does not exist here (probed:
no matlab, octave, oct2py or matlab.engine in the image or the wheelhouse), and nothing in the repository stores a Darcy
output, so parity with utilities/mfiles/*.m stays UNPINNED.  These tests tie both restatements -- the scipy oracle
(oracle/darcy_oracle.py) and the host-assembled operators the CUDA path multiplies with (ces_b200/darcy.py) -- to
closed forms instead of to each other:

  * theta = 0  =>  -Laplace(p) = 1 on the unit square, p = 0 on the boundary: the double sine series, second-order
    convergence of the 5-point scheme (solve_gwf.m:16-35);
  * one KL coefficient  =>  theta = scale * cos(pi k1 x) cos(pi k2 y) at the cell centres (gaussrnd_coarse.m:9-21 with
    idct2 = orthonormal inverse DCT-II);
  * the not-a-knot spline (interp2 'spline', solve_gwf.m:13,37) reproduces cubics exactly, extrapolation included.
The GPU counterpart of the first check is tests/test_gpu_darcy.py::test_poisson_limit_matches_the_series_solution.
"""
import numpy as np
import pytest

from ces_b200 import darcy as cdarcy
from oracle import darcy_oracle as do


def poisson_series(x, y, terms=399):
    """p(x, y) = sum_{m, n odd} 16 / (pi^4 m n (m^2 + n^2)) sin(m pi x) sin(n pi y): -Laplace(p) = 1, p = 0 on the boundary."""
    m = np.arange(1, terms + 1, 2)
    coef = 16.0 / (np.pi ** 4 * np.outer(m, m) * (m[:, None] ** 2 + m[None, :] ** 2))
    return np.sin(np.pi * np.outer(x, m)) @ coef @ np.sin(np.pi * np.outer(m, y))


def test_oracle_poisson_limit_second_order():
    errs = []
    for K in (16, 32, 64):
        centres = (np.arange(K) + 0.5) / K
        p = do.solve_gwf(np.zeros((K, K)))
        exact = poisson_series(centres, centres)
        errs.append(np.abs(p - exact).max())
        assert errs[-1] < 0.6 / (K - 1) ** 2                      # O(h^2), constant from the K = 16 run
    assert abs(p.max() - 0.0736713) < 2e-4                        # max of the torsion function of the unit square
    assert 3.0 < errs[0] / errs[1] < 5.5 and 3.0 < errs[1] / errs[2] < 5.5


@pytest.mark.parametrize("N,k1,k2", [(16, 1, 0), (16, 0, 3), (32, 2, 5), (48, 7, 7)])
def test_single_kl_mode_is_a_cosine(N, k1, k2):
    alpha, tau = 2.0, 3.0
    xi = np.zeros((N, N))
    xi[k1, k2] = 1.3
    centres = (np.arange(N) + 0.5) / N
    amp = 1.3 * tau ** (alpha - 1) * (np.pi ** 2 * (k1 ** 2 + k2 ** 2) + tau ** 2) ** (-alpha / 2)
    amp *= (np.sqrt(2.0) if k1 else 1.0) * (np.sqrt(2.0) if k2 else 1.0)        # N * w_k1 * w_k2 of the orthonormal DCT
    exact = amp * np.outer(np.cos(np.pi * k1 * centres), np.cos(np.pi * k2 * centres))
    assert np.allclose(do.gaussrnd_coarse(xi, alpha, tau, N), exact, rtol=0, atol=1e-14)
    # the operator the device multiplies with (Theta = U^T Phi^T)
    Phi = cdarcy.kl_operator(N, alpha, tau, [k1 * N + k2])
    assert np.allclose(1.3 * Phi[0].reshape(N, N), exact, rtol=0, atol=1e-14)
    # the constant mode is removed (gaussrnd_coarse.m:19)
    assert np.all(cdarcy.kl_operator(N, alpha, tau, [0]) == 0.0)


@pytest.mark.parametrize("K", [8, 21, 64])
def test_not_a_knot_spline_is_exact_on_cubics(K):
    centres = (np.arange(K) + 0.5) / K
    nodes = np.arange(K) / (K - 1.0)                               # first / last node lie half a cell outside the sites
    f = lambda x: 0.3 - 1.1 * x + 2.0 * x ** 2 - 0.7 * x ** 3
    for sites, query in ((centres, nodes), (nodes, centres)):
        S = cdarcy.spline_operator(sites, query)
        assert np.allclose(S @ f(sites), f(query), rtol=0, atol=1e-12)
        V = np.outer(f(sites), 1.0 + sites)                        # the oracle's separable interp2
        assert np.allclose(do._interp2_spline(sites, V, query), np.outer(f(query), 1.0 + query), rtol=0, atol=1e-12)
        assert np.allclose(S.sum(axis=1), 1.0, atol=1e-12)         # constants are reproduced: theta = 0 gives c = 1 exactly


def test_oracle_solution_is_symmetric_for_a_symmetric_field():
    """xi symmetric under transposition => theta(x, y) = theta(y, x) => p(x, y) = p(y, x): catches a transposed
    assembly (the spdiags sub/super-diagonal offsets and the vec2mat + ' pair of solve_gwf.m:18-37)."""
    N = 24
    rng = np.random.default_rng(0)
    xi = rng.standard_normal((N, N))
    xi = 0.5 * (xi + xi.T)
    p = do.solve_gwf(do.gaussrnd_coarse(xi, 2.0, 3.0, N))
    assert np.allclose(p, p.T, rtol=1e-10, atol=1e-14)
    # and an asymmetric field gives an asymmetric solution, transposing with the field
    xi2 = rng.standard_normal((N, N))
    p2 = do.solve_gwf(do.gaussrnd_coarse(xi2, 2.0, 3.0, N))
    p2t = do.solve_gwf(do.gaussrnd_coarse(xi2.T, 2.0, 3.0, N))
    assert np.allclose(p2t, p2.T, rtol=1e-9, atol=1e-13) and not np.allclose(p2, p2.T, rtol=1e-3)
