"""Load the *real* reference (agarbuno/ces) for oracle validation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``ces/calibrate.py`` mixes tabs and spaces inside ``class sampling``
(ces/calibrate.py:243-247, 362-369, 451, 492-496) so ``import ces.calibrate``
raises ``TabError`` on Python 3.  The file is read, tab-expanded to 4 columns in
memory and exec'd into a private module; nothing under /root/reference is
modified.  ``ces.utils`` imports cleanly and is loaded through importlib.

The reference itself only exists in the build container (/root/reference).  ``oracle/stage_reference.py`` copies the
three files of the update path, unmodified, to the git-ignored ``baseline/_ref/`` so that they travel to the GPU box
for ``bench.py --impl reference``; this loader looks at /root/reference (or ``CES_REFERENCE_ROOT``) first and there
second.  Tests must not depend on either: on the GPU box they use the committed golden vectors.
"""
import importlib.util
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _find_root():
    for root in (os.environ.get("CES_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if root and os.path.isfile(os.path.join(root, "ces", "calibrate.py")):
            return root
    return os.environ.get("CES_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()

_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ces", "calibrate.py"))


def load_calibrate():
    """Return a module object holding the reference ``enka`` and ``sampling``."""
    if "calibrate" in _cache:
        return _cache["calibrate"]
    path = os.path.join(REFERENCE_ROOT, "ces", "calibrate.py")
    with open(path) as fh:
        src = fh.read().expandtabs(4)
    mod = types.ModuleType("ces_calibrate_reference")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache["calibrate"] = mod
    return mod


def load_utils():
    """Return the reference ``ces.utils`` module (lineal, elliptic, banana ...)."""
    if "utils" in _cache:
        return _cache["utils"]
    path = os.path.join(REFERENCE_ROOT, "ces", "utils.py")
    spec = importlib.util.spec_from_file_location("ces_utils_reference", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["utils"] = mod
    return mod


def load_sample():
    """The reference ``ces/sample.py`` (class ``MCMC``).  Its module-level ``import gpflow`` and the package-relative
    imports of ``calibrate`` / ``emulate`` cannot be satisfied here (GPflow is absent, calibrate has the TabError), so the
    file is exec'd, unmodified, into a private module with stand-ins for those three names: ``model_mh`` -- the only
    method used -- touches none of them.  Needs ``ces/sample.py`` under REFERENCE_ROOT (not staged for the GPU box: the
    GPU tests use golden vectors made here)."""
    if "sample" in _cache:
        return _cache["sample"]
    path = os.path.join(REFERENCE_ROOT, "ces", "sample.py")
    with open(path) as fh:
        src = fh.read().expandtabs(4)
    src = src.replace("from . import calibrate", "calibrate = None").replace("from . import emulate", "emulate = None")
    src = src.replace("import gpflow as gp", "gp = None")
    src = src.replace("from tqdm.autonotebook import tqdm", "from tqdm import tqdm")
    mod = types.ModuleType("ces_sample_reference")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache["sample"] = mod
    return mod


def make_sampler(p, n_obs, J, mu, sigma, ustar, T=30, t_hist=None):
    """A reference ``sampling`` object primed so one update rule can be called
    directly, outside ``run`` (which is what creates these attributes,
    ces/calibrate.py:306-339).  ``t_hist`` is the list of cumulative times of the
    steps already taken; the reference detects "first step" by
    ``len(self.Uall) == 1`` (ces/calibrate.py:262)."""
    import numpy as np

    mod = load_calibrate()
    eks = mod.sampling(p=p, n_obs=n_obs, J=J)
    eks.mu = np.asarray(mu, dtype=float).reshape(p, -1)
    eks.sigma = np.asarray(sigma, dtype=float)
    eks.ustar = np.asarray(ustar, dtype=float).reshape(p, -1)
    eks.T = T
    t_hist = list(t_hist or [])
    eks.Uall = [None] * (len(t_hist) + 1)
    eks.radspec = []
    eks.metrics = {"self-bias": [], "bias": [], "self-bias-data": [], "bias-data": [], "t": t_hist}
    return eks


def reference_step(rule, y, U, G, Gamma, mu, sigma, ustar, xi, T=30, t_hist=None, **kwargs):
    """Run one reference update on the given inputs with the given noise.

    The reference draws its noise from the global numpy RNG
    (``np.random.normal(0, 1, [p, J])``, ces/calibrate.py:447,488,527).  To make it
    consume exactly ``xi`` the global ``np.random.normal`` is swapped for a stub
    for the duration of the call (the reference file itself is untouched).
    Returns ``(Uk, hk, metrics_dict)``.
    """
    import numpy as np

    p, J = U.shape
    eks = make_sampler(p, G.shape[0], J, mu, sigma, ustar, T=T, t_hist=t_hist)
    fn = {"eks": eks.eks_update, "aldi": eks.eks_update_aldi,
          "aldi_constant": eks.eks_update_aldi_constant}[rule]
    t_before = eks.metrics["t"][-1] if eks.metrics["t"] else 0.0
    n_before = len(eks.metrics["t"])
    real_normal = np.random.normal

    def fake_normal(loc=0.0, scale=1.0, size=None):
        assert list(size) == [p, J], size
        return loc + scale * np.array(xi, dtype=float, copy=True)

    seen = {}
    inner = eks.timestep_method

    def recording_timestep(*a, **kw):
        seen["hk"] = inner(*a, **kw)
        return seen["hk"]

    eks.timestep_method = recording_timestep
    np.random.normal = fake_normal
    try:
        Uk = fn(y, U, G, Gamma, 0, **kwargs)
    finally:
        np.random.normal = real_normal
    t_after = eks.metrics["t"][-1]
    # aldi_constant computes hk inline (ces/calibrate.py:519-523); recover it from
    # the time bookkeeping (exact when this is the first step).
    hk = seen.get("hk", t_after - (t_before if n_before else 0.0))
    m = {key: eks.metrics[key][-1] for key in ("self-bias", "bias", "self-bias-data", "bias-data")}
    m["t"] = t_after
    return Uk, hk, m
