"""oracle/mcmc_oracle.py against the chains of the REAL reference (tests/golden/mcmc_cases.npz): the numpy restatement of
MCMC.model_mh is pinned -- samples, acceptance rate and the generator state after the run -- for every golden case,
including the resumed second call.  CPU only."""
import os

import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import forward_oracle as fo, mcmc_oracle as mo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcmc_cases.npz"))


def _forward(name):
    kind = str(G["%s/kind" % name])
    if kind == "lineal":
        A = G["%s/A" % name]
        return (lambda th: fo.lineal(A, np.asarray(th).reshape(-1, 1))[:, 0]), 0
    if kind == "elliptic":
        return (lambda th: fo.elliptic(np.asarray(th).reshape(2, 1))[:, 0]), 0
    return (lambda th: fo.banana(np.asarray(th).reshape(2, 1))[:, 0]), 2          # banana draws two normals per evaluation


@pytest.mark.parametrize("name", [str(n) for n in G["names"]])
def test_oracle_chain_is_the_reference_chain(name):
    forward, draws = _forward(name)
    prior = multivariate_normal(G["%s/prior_mean" % name], G["%s/prior_cov" % name])
    kw = dict(delta=float(G["%s/delta" % name]), enka_scaling=bool(G["%s/scaling" % name]), forward_draws=draws)
    if str(G["%s/update" % name]):
        kw.update(update=str(G["%s/update" % name]), beta=float(G["%s/beta" % name]))
    np.random.seed(13)
    samples, accept = mo.model_mh(forward, int(G["%s/n" % name]), prior, G["%s/Ustar" % name], G["%s/y" % name], G["%s/Gamma" % name], **kw)
    want = G["%s/samples" % name]
    assert samples.shape == want.shape and np.abs(samples - want).max() <= 1e-12 * np.abs(want).max()
    assert accept == float(G["%s/accept" % name])
    assert np.array_equal(np.random.get_state()[1], G["%s/state1_key" % name])
    samples2, accept2 = mo.model_mh(forward, 50, prior, G["%s/Ustar" % name], G["%s/y" % name], G["%s/Gamma" % name], samples=samples, **kw)
    want2 = G["%s/samples_resumed" % name]
    assert samples2.shape == want2.shape and np.abs(samples2 - want2).max() <= 1e-12 * np.abs(want2).max()
    assert accept2 == float(G["%s/accept_resumed" % name]) and np.array_equal(np.random.get_state()[1], G["%s/state2_key" % name])
