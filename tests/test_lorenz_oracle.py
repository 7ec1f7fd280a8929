"""oracle/lorenz_oracle.py and the host-side pieces of ces_b200.utils' Lorenz classes against golden vectors made with
the real reference classes (tests/golden/make_golden_lorenz.py; ces/utils.py:124-447).  CPU only."""
import os

import numpy as np
import pytest

from oracle import lorenz_oracle as lo

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lorenz_cases.npz"))


def test_l63_oracle_tracks_the_reference_integrator_over_a_short_horizon():
    """RK4 with 10 steps per output interval vs scipy odeint (rtol ~1.5e-8) over T = 2: same trajectory to 1e-3 of its
    range despite the chaos, same window statistics to 1e-4 relative."""
    t = GOLD["l63_t"]
    for i in range(3):
        ws = lo.l63_solve(GOLD["l63_w0"][i], GOLD["l63_args"][i], len(t), t[1] - t[0], 10)
        assert np.abs(ws - GOLD["l63_ws"][i]).max() < 1e-3
        st = lo.l63_statistics(ws, 100)
        assert np.abs(st - GOLD["l63_stats"][i]).max() < 1e-4 * np.abs(GOLD["l63_stats"][i]).max()
        wl = lo.l63_solve(GOLD["l63_w0"][i], np.log(GOLD["l63_args"][i]), len(t), t[1] - t[0], 10, log_params=True)
        assert np.abs(wl - GOLD["l63log_ws"][i]).max() < 1e-3
        # statistics of the reference's own trajectory: exact
        assert np.abs(lo.l63_statistics(GOLD["l63_ws"][i], 100) - GOLD["l63_stats"][i]).max() < 1e-12


@pytest.mark.parametrize("name,keys", [("full", ("h", "F", "log_c", "b")), ("Fc", ("F", "log_c")), ("Fb", ("F", "b")),
                                       ("hFb", ("h", "F", "b")), ("hcb", ("h", "log_c", "b"))])
def test_l96_right_hand_side_matches_reference(name, keys):
    w = GOLD["l96_state"]
    par = dict(zip(keys, GOLD["l96_args_" + name]))
    got = lo.l96_rhs(w, 6, 4, par.get("h", 1.0), par.get("F", 10.0), np.exp(par.get("log_c", np.log(10.0))), par.get("b", 10.0))
    assert np.abs(got - GOLD["l96_rhs_" + name]).max() < 1e-13 * np.abs(GOLD["l96_rhs_" + name]).max()


def test_l96_statistics_match_reference():
    assert np.abs(lo.l96_statistics(GOLD["l96_traj"], 6, 4, 11, 10) - GOLD["l96_stats"]).max() < 1e-13
    phi = lo.l96_statistics(GOLD["l96hom_traj"], 8, 4, 11, 10).reshape(5, -1)
    assert np.abs(phi.mean(axis=1) - GOLD["l96hom_stats"]).max() < 1e-13
    assert np.abs(phi[:, 7] - GOLD["l96hom_stats_k7"]).max() < 1e-13


def test_host_side_of_the_product_classes():
    """Constructors, names, the single-state right-hand sides and ``statistics`` of ces_b200.utils' Lorenz classes
    (pure host logic; the integration itself only exists on the device)."""
    from ces_b200 import utils as cu

    m = cu.lorenz63(l_window=1, freq=100)
    assert (m.n_state, m.n_obs, m.type, m.model_name, repr(m), str(m)) == (3, 9, 'pde', 'lorenz63', 'lorenz63', 'lorenz633')
    assert np.abs(m.statistics(GOLD["l63_ws"][0]) - GOLD["l63_stats"][0]).max() < 1e-12
    assert np.allclose(m([1.0, 2.0, 3.0], 0.0, 28.0, 2.0), lo.l63_rhs(np.array([1.0, 2.0, 3.0]), 10.0, 28.0, 2.0))
    ml = cu.lorenz63_log()
    assert ml.model_name == 'lorenz63_log'
    assert np.allclose(ml([1.0, 2.0, 3.0], 0.0, np.log(28.0), np.log(2.0)), lo.l63_rhs(np.array([1.0, 2.0, 3.0]), 10.0, 28.0, 2.0))
    for name, cls in (("full", cu.lorenz96), ("Fc", cu.lorenz96Fc), ("Fb", cu.lorenz96Fb), ("hFb", cu.lorenz96hFb),
                      ("hcb", cu.lorenz96hcb)):
        mod = cls(n_slow=6, n_fast=4) if cls is cu.lorenz96 else cls()
        mod.n_slow, mod.n_fast, mod.n_state = 6, 4, 30
        got = mod(0.0, GOLD["l96_state"], *GOLD["l96_args_" + name])
        assert np.abs(got - GOLD["l96_rhs_" + name]).max() < 1e-13 * np.abs(GOLD["l96_rhs_" + name]).max(), name
    mod = cu.lorenz96(n_slow=6, n_fast=4, l_window=1, freq=10, spinup=1)
    assert np.abs(mod.statistics(GOLD["l96_traj"]) - GOLD["l96_stats"]).max() < 1e-13
    hom = cu.lorenz96_hom()
    hom.n_slow, hom.n_fast, hom.n_state, hom.l_window, hom.freq, hom.spinup = 8, 4, 40, 1, 10, 1
    assert np.abs(hom.statistics(GOLD["l96hom_traj"]) - GOLD["l96hom_stats"]).max() < 1e-13
    hom.hom = False
    assert np.abs(hom.statistics(GOLD["l96hom_traj"]) - GOLD["l96hom_stats_k7"]).max() < 1e-13
    assert repr(cu.lorenz96Fc()) == 'lorenz96,36,10,2' and repr(cu.lorenz96()) == 'lorenz96,36,10'
    with pytest.raises(ValueError):
        m._grid(np.array([0.0, 0.1, 0.3]))
