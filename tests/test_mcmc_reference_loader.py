"""The golden MCMC chains (tests/golden/mcmc_cases.npz) are what the REAL reference produces today: re-run one case through
oracle/reference_loader.load_sample() (ces/sample.py exec'd unmodified with stand-ins for its unavailable imports) and
compare.  Needs /root/reference (build container)."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import reference_loader as rl

import os  # noqa: E402

pytestmark = pytest.mark.skipif(not (rl.available() and os.path.isfile(os.path.join(rl.REFERENCE_ROOT, "ces", "sample.py"))),
                                reason="ces/sample.py of the reference is only present in the build container "
                                       "(baseline/_ref stages calibrate.py and utils.py only)")


def test_golden_chain_is_the_live_reference():
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcmc_cases.npz"))
    cal, utils, sample = rl.load_calibrate(), rl.load_utils(), rl.load_sample()
    name = "lineal3_dense"
    Ustar, y = g["%s/Ustar" % name], g["%s/y" % name]
    enka = cal.sampling(p=Ustar.shape[0], n_obs=y.shape[0], J=Ustar.shape[1])
    enka.Ustar = Ustar
    mc = sample.MCMC()
    mc.mute_bar, mc.y_obs = True, y
    np.random.seed(13)
    mc.model_mh(utils.lineal(g["%s/A" % name]), int(g["%s/n" % name]), multivariate_normal(g["%s/prior_mean" % name], g["%s/prior_cov" % name]),
                enka, g["%s/Gamma" % name], delta=float(g["%s/delta" % name]), enka_scaling=bool(g["%s/scaling" % name]))
    assert np.array_equal(mc.samples, g["%s/samples" % name]) and mc.accept == float(g["%s/accept" % name])
    assert np.array_equal(np.random.get_state()[1], g["%s/state1_key" % name])
