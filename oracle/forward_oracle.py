"""numpy restatement of the map-type forward models of ces/utils.py, batched
over the ensemble (one column per particle), noise-free (``flag_noise=False``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity PINNED against the
reference classes evaluated particle by particle through ``enka.G_ens``
(ces/calibrate.py:106-130) in tests/test_oracle.py.
"""
import numpy as np


def lineal(A, U, b=0.0):
    """ces/utils.py:25-31: A theta + b for every column."""
    return A @ U + (np.asarray(b, dtype=float).reshape(-1, 1) if np.ndim(b) else b)


def lineal_log(A, U, b=0.0):
    """ces/utils.py:39-42: A exp(phi) (+ b, always 0 in the reference constructor :34)."""
    return lineal(A, np.exp(U), b)


def elliptic(U, x1=0.25, x2=0.75):
    """ces/utils.py:72-89: p(x) = u2 x + exp(-u1) (x - x^2)/2 at x1, x2."""
    u1, u2 = U[0], U[1]
    e = np.exp(-u1)
    return np.stack([u2 * x1 + e * (-x1 ** 2 + x1) * 0.5,
                     u2 * x2 + e * (-x2 ** 2 + x2) * 0.5])


def banana(U, a=1.0, b=0.5):
    """ces/utils.py:116-122: (a u1, u2/a - b (u1^2 + a^2))."""
    u1, u2 = U[0], U[1]
    return np.stack([u1 * a, u2 / a - b * (u1 ** 2 + a ** 2)])
