"""The restatement of the device's multilevel-preconditioned CG (oracle/darcy_pcg_oracle.py) against the sparse direct
solve of the Darcy oracle: same nodal pressure, and the iteration counts the preconditioner levels buy.  CPU only."""
import numpy as np
import pytest

from oracle import darcy_oracle as do, darcy_pcg_oracle as dp


def _direct_nodal(theta):
    import scipy.sparse.linalg as spla

    K = theta.shape[0]
    A, n = dp.system(theta)
    x = spla.spsolve(A.tocsc(), np.full(n * n, 1.0 / (K - 1) ** 2))
    P = np.zeros((K, K))
    P[1:K - 1, 1:K - 1] = x.reshape(n, n)
    return P


@pytest.mark.parametrize("N,p,scale", [(32, 24, 1.0), (64, 64, 1.0), (64, 64, 10.0)])
def test_multilevel_pcg_matches_direct_solve(N, p, scale):
    m = do.ModelTrunc(Nmesh=N, p=p)
    theta = m.eval_rf(scale * np.random.default_rng(N).standard_normal(p))
    want = _direct_nodal(theta)
    counts = {}
    for levels in ((), ("coarse",), ("pair", "coarse")):
        got, counts[levels] = dp.solve(theta, levels=levels)
        assert np.abs(got - want).max() < 1e-10 * np.abs(want).max(), levels
    # each level pays: Jacobi alone > + coarse level > + 4 x 4 level (which coincides with the coarse level at Nmesh = 32)
    assert counts[()] > 1.8 * counts[("coarse",)]
    if dp.coarse_size(N) > 4:
        assert counts[("pair", "coarse")] < 0.9 * counts[("coarse",)]
    else:
        assert counts[("pair", "coarse")] == counts[("coarse",)]


def test_system_is_the_oracles_matrix():
    """dp.system is the matrix of darcy_oracle.solve_gwf (row-major instead of column-block ordering, without (K-1)^2)."""
    m = do.ModelTrunc(Nmesh=16, p=10)
    theta = m.eval_rf(np.random.default_rng(0).standard_normal(10))
    P = _direct_nodal(theta)
    centres = (np.arange(16) + 0.5) / 16
    nodes = np.arange(16) / 15.0
    back = do._interp2_spline(nodes, P, centres)
    assert np.abs(back - do.solve_gwf(theta)).max() < 1e-12 * np.abs(back).max()
