"""``MCMC`` with the interface of ``ces/sample.py``: the Sample stage that consumes a calibrated ensemble.

``model_mh`` -- random-walk Metropolis-Hastings / pCN on the TRUE forward model (ces/sample.py:121-196) -- runs whole
chains on the GPU (``ces_mcmc_model_mh``, csrc/mcmc.cu): one warp per chain, the proposal / forward map / misfit /
accept loop never leaves the kernel.  Same call signature, same results attributes (``samples`` (p, n + 1), ``accept``),
same resume behaviour (a second call continues from the last sample) and -- for the chain the reference would run -- the
same random numbers: the kernel consumes numpy's global MT19937 stream exactly as ``np.random.normal(0, 1, p)`` followed
by ``np.random.uniform()`` would, and leaves the generator where the reference's loop would leave it.  ``n_chains > 1``
(a new, additive kwarg) runs further independent chains beside it from device noise: ``samples_chains`` /
``accept_chains``.

The forward model must be one of the device maps of ``ces_b200.utils`` (``lineal``, ``lineal_log``, ``elliptic``,
``banana``) with p <= 32, k <= 64.  ``gp_mh`` (Metropolis-Hastings on a GPflow emulator, ces/sample.py:17-119) is not
provided: it needs trained GPflow models (``emulate.predict_gps``, ces/emulate.py:17-79), a third-party library that is
absent here and whose predictions could not be pinned; that half of SURVEY.md section 8f-4 stays out of scope.
"""
from __future__ import print_function

import ctypes

import numpy as np

from . import _lib


class MCMC(object):

    def __init__(self):
        self.mute_bar = False

    def gp_mh(self, enka, n_mcmc, prior, delta=1., enka_scaling=True, **kwargs):
        raise NotImplementedError("gp_mh samples a GPflow emulator (ces/sample.py:17-119, ces/emulate.py:17-79); GPflow is "
                                  "not part of this package -- use model_mh on the forward model itself")

    def model_mh(self, model, n_mcmc, prior, enka, Gamma, delta=1., enka_scaling=True, **kwargs):
        """ces/sample.py:121-196.  ``prior``: an object with ``logpdf`` / ``mean`` / ``cov`` (scipy.stats
        multivariate_normal, as in the reference's notebooks).  kwargs: ``update='pCN'``, ``beta`` (reference);
        ``n_chains``, ``seed`` (device chains beside the reference's one)."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("ces_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        kind = getattr(model, 'device_kind', None)
        if getattr(model, 'type', None) != 'map' or kind not in _lib.MAPS or getattr(model, 'flag_noise', False):
            raise NotImplementedError("model_mh on the device covers the ces_b200.utils maps (lineal, lineal_log, elliptic, "
                                      "banana) without in-model noise; got %r" % (model,))
        p, k = int(enka.p), int(enka.n_obs)
        pcn = kwargs.get('update', None) == 'pCN'
        if kwargs.get('update', None) not in (None, 'pCN'):
            raise ValueError("unknown update %r" % (kwargs.get('update'),))     # the reference would fail on `proposal`
        # RW needs the proposal distribution (:123-129); pCN proposes according to the prior
        if enka_scaling:
            scales = delta * np.linalg.cholesky(np.cov(enka.Ustar).reshape(p, p))
        else:
            scales = delta * np.eye(p)
        if pcn:
            scales = np.linalg.cholesky(np.asarray(prior.cov, dtype=np.float64).reshape(p, p))
        mean = np.asarray(enka.Ustar, dtype=np.float64).mean(axis=1)            # :131
        y = np.ascontiguousarray(np.asarray(self.y_obs, dtype=np.float64).reshape(k))
        Gam = np.asarray(Gamma, dtype=np.float64).reshape(k, k)
        Ginv2 = np.ascontiguousarray(np.linalg.inv(2.0 * Gam))
        mu = Pinv = None
        if not pcn:
            mu = np.ascontiguousarray(np.asarray(prior.mean, dtype=np.float64).reshape(p))
            Pinv = np.ascontiguousarray(np.linalg.inv(np.asarray(prior.cov, dtype=np.float64).reshape(p, p)))
        # resume (:156-164): continue from the last sample; Phi_current stays the ensemble mean's, like the reference
        previous = None
        if hasattr(self, 'samples'):
            previous = np.asarray(self.samples, dtype=np.float64)
            first = previous[:, -1].copy()
        else:
            first = mean.copy()
        n_chains = int(kwargs.get('n_chains', 1))
        seed = int(kwargs.get('seed', 0))
        start = np.tile(first, (n_chains, 1))
        if n_chains > 1:
            # the extra chains start from ensemble members (over-dispersed starts), deterministically chosen
            J = enka.Ustar.shape[1]
            start[1:] = np.asarray(enka.Ustar, dtype=np.float64).T[np.arange(1, n_chains) % J]
        phi_point = np.tile(mean, (n_chains, 1))
        phi_point[1:] = start[1:]
        # numpy's global generator: ('MT19937', key[624], pos, has_gauss, cached_gaussian)
        st = np.random.get_state()
        mt = np.empty(626, dtype=np.uint32)
        mt[:624], mt[624], mt[625] = st[1], st[2], st[3]
        gauss = np.array([st[4]], dtype=np.float64)
        samples = np.empty((n_chains, int(n_mcmc) + 1, p))
        accepted = np.zeros(n_chains, dtype=np.int32)
        A_dev, lda, b_dev, params = model._device_args(torch)
        par = np.ascontiguousarray(params, dtype=np.float64) if params is not None else None
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().ces_mcmc_model_mh(
            ctypes.c_void_p(stream), _lib.MAPS[kind], p, k,
            ctypes.c_void_p(A_dev.data_ptr()) if A_dev is not None else None, int(lda),
            ctypes.c_void_p(b_dev.data_ptr()) if b_dev is not None else None,
            _lib.host_ptr(par) if par is not None else None, _lib.host_ptr(y), _lib.host_ptr(Ginv2),
            _lib.host_ptr(mu) if mu is not None else None, _lib.host_ptr(Pinv) if Pinv is not None else None,
            _lib.host_ptr(np.ascontiguousarray(scales, dtype=np.float64)), 1 if pcn else 0, float(kwargs.get('beta', 0.5)),
            int(n_mcmc), n_chains, _lib.host_ptr(start), _lib.host_ptr(phi_point), mt.ctypes.data, _lib.host_ptr(gauss), seed,
            _lib.host_ptr(samples), accepted.ctypes.data))
        np.random.set_state(('MT19937', mt[:624].copy(), int(mt[624]), int(mt[625]), float(gauss[0])))
        chain0 = samples[0]
        if previous is not None:
            # the reference appends to list(self.samples.T) WITHOUT repeating the state it resumes from (:158-160)
            chain0 = np.vstack([previous.T, samples[0][1:]])
        self.samples = np.array(chain0).T                    # (p, n + 1), :195
        self.accept = accepted[0] / float(n_mcmc)            # :196
        if n_chains > 1:
            self.samples_chains = samples.transpose(0, 2, 1)     # (n_chains, p, n + 1)
            self.accept_chains = accepted / float(n_mcmc)

    def random_walk(self, current, scales, n_dim):
        """ces/sample.py:198-199 (host helper, kept for API parity; the device chains draw their own)."""
        return current + np.matmul(scales, np.random.normal(0, 1, n_dim))

    def pCN(self, current, scales, n_dim, beta=0.5):
        """ces/sample.py:201-202."""
        return np.sqrt(1 - beta ** 2) * current + np.sqrt(beta) * np.matmul(scales, np.random.normal(0, 1, n_dim))
