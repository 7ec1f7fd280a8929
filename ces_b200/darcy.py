"""2-D Darcy flow forward model with the interface of ``ces/darcy.py`` (``model``, ``model_trunc``),
solved for a whole ensemble on the GPU (one cluster of CTAs per member).

Reference: ``ces/darcy.py:9-138`` drives a MATLAB engine that runs
``utilities/mfiles/gaussrnd_coarse.m`` (KL coefficients -> log-permeability by inverse 2-D DCT) and
``utilities/mfiles/solve_gwf.m`` (spline to the nodes, 5-point finite-difference system with
arithmetic-mean face coefficients, sparse direct solve, spline back to the cell centres) once per
particle.  Here ``start``/``stop``/``set_rnd_seed`` are kept as no-ops (there is no engine) and the
arithmetic runs in libces_b200.so (``ces_darcy_*``):

  1. Theta = U^T Phi^T            one DMMA GEMM; Phi holds the scaled 2-D DCT basis of the active KL modes
  2. a = exp(Theta)               cell-centre permeability
  3. c = S a S^T                  not-a-knot spline centres -> nodes as two DMMA GEMMs (stacked / batched)
  4. -div(c grad p) = 1           multilevel-preconditioned CG, member resident in registers + shared memory of a cluster
  5. P = S2 p S2^T                spline nodes -> centres (two more GEMMs), gather at ``obs_index``

The constant operators Phi, S, S2 depend only on (Nmesh, alpha, tau, p) and are assembled once on the
host below.  Parity with the MATLAB original is unpinned (no MATLAB anywhere): the tests
compare against the scipy restatement in ``oracle/darcy_oracle.py``.
"""
import ctypes

import numpy as np

from . import _lib


# ------------------------------------------------------------------------------------------------
# constant operators (host, one-time)
def kl_scale(N, alpha, tau):
    """N * tau^(alpha-1) (pi^2 (k1^2 + k2^2) + tau^2)^(-alpha/2), zero for the constant mode
    (gaussrnd_coarse.m:15-19)."""
    k = np.arange(N)
    K1, K2 = np.meshgrid(k, k)
    s = N * (tau ** (alpha - 1)) * (np.pi ** 2 * (K1 ** 2 + K2 ** 2) + tau ** 2) ** (-alpha / 2)
    s[0, 0] = 0.0
    return s


def mode_rank(N, alpha, tau):
    """ces/darcy.py:74-82."""
    k = np.arange(N)
    K1, K2 = np.meshgrid(k, k)
    eigs = (tau ** (alpha - 1)) * (np.pi ** 2 * (K1 ** 2 + K2 ** 2) + tau ** 2) ** (-alpha / 2)
    eigs[0, 0] = 1e-10
    return (-eigs).flatten().argsort()


def idct_basis(N):
    """B[n, k] = w_k cos(pi (2n+1) k / (2N)), w_0 = 1/sqrt(N), w_k = sqrt(2/N): the orthonormal inverse
    DCT-II, so idct2(L) = B L B^T."""
    n = np.arange(N)[:, None]
    k = np.arange(N)[None, :]
    B = np.cos(np.pi * (2 * n + 1) * k / (2.0 * N)) * np.sqrt(2.0 / N)
    B[:, 0] = 1.0 / np.sqrt(N)
    return B


def kl_operator(N, alpha, tau, modes):
    """Phi^T (len(modes) x N^2): row m is the field of unit coefficient on flat mode index modes[m]
    (row-major reshape of xi to N x N, ces/darcy.py:92,138)."""
    B = idct_basis(N)
    s = kl_scale(N, alpha, tau)
    out = np.empty((len(modes), N * N))
    for r, m in enumerate(modes):
        k1, k2 = divmod(int(m), N)
        out[r] = (s[k1, k2] * np.outer(B[:, k1], B[:, k2])).reshape(-1)
    return out


def spline_operator(x, xq):
    """Dense matrix of not-a-knot cubic-spline interpolation from uniform sites x to queries xq
    (queries outside [x0, x_end] use the end polynomials, as MATLAB's spline / interp2 'spline')."""
    x = np.asarray(x, dtype=float)
    n = x.shape[0]
    if n < 4:
        raise ValueError("not-a-knot splines need at least 4 sites")
    h = x[1] - x[0]
    # second derivatives M = T^-1 R y
    T = np.zeros((n, n))
    R = np.zeros((n, n))
    T[0, 0:3] = [1.0, -2.0, 1.0]
    T[n - 1, n - 3:n] = [1.0, -2.0, 1.0]
    for i in range(1, n - 1):
        T[i, i - 1:i + 2] = [1.0, 4.0, 1.0]
        R[i, i - 1:i + 2] = np.array([1.0, -2.0, 1.0]) * (6.0 / h ** 2)
    Mop = np.linalg.solve(T, R)
    S = np.zeros((len(xq), n))
    eye = np.eye(n)
    for q, xv in enumerate(xq):
        i = int(np.clip(np.floor((xv - x[0]) / h), 0, n - 2))
        t = xv - x[i]
        # s = y_i + t[(y_{i+1}-y_i)/h - h(2M_i + M_{i+1})/6] + t^2 M_i/2 + t^3 (M_{i+1}-M_i)/(6h)
        S[q] = (eye[i] + t * ((eye[i + 1] - eye[i]) / h - h * (2 * Mop[i] + Mop[i + 1]) / 6.0)
                + t ** 2 * Mop[i] / 2.0 + t ** 3 * (Mop[i + 1] - Mop[i]) / (6.0 * h))
    return S


# ------------------------------------------------------------------------------------------------
class model(object):
    """Darcy flow with all N^2 KL coefficients as parameters.  ces/darcy.py:9-98."""
    device_kind = "darcy"

    def __init__(self, alpha=2., tau=3., Nmesh=2. ** 4):
        self.alpha = alpha
        self.tau = tau
        self.Nmesh = Nmesh
        self.p = int(self.Nmesh * self.Nmesh)
        self.model_name = 'darcy-flow'
        self.type = 'map'
        self.flag_noise = False
        # Relative residual (in the preconditioner's norm) at which the CG solve stops; the reference solves directly.
        # Measured on 64^2 / 96^2 / 128^2 fields at prior scales 1-10 (tools/darcy_tol_experiment.py): the error against the
        # sparse direct solve is 0.2-0.35 x tol, uniformly.  1e-10 keeps the field within 3e-11 of the direct solve -- 30 x
        # inside the 1e-9 parity bound of tests/test_gpu_darcy.py -- for 22 % fewer iterations than 1e-13 (set it back to
        # 1e-13 for direct-solver accuracy).
        self.tol = 1e-10
        self.max_iter = 0           # 0: the library default, 40 N iterations
        self._dev = None

    # engine management of the reference: nothing to start here
    def start(self, mpath=None):
        return None

    def stop(self):
        self._release()

    def set_rnd_seed(self, seed=1):
        return None

    def set_initial(self, seed=1):
        np.random.seed(seed)
        self.ustar = np.random.normal(0, 1, int(self.p))

    def set_rank(self):
        N = int(self.Nmesh)
        k = np.arange(N)
        K1, K2 = np.meshgrid(k, k)
        self.eigs = (self.tau ** (self.alpha - 1)) * (np.pi ** 2 * (K1 ** 2 + K2 ** 2) + self.tau ** 2) ** (-self.alpha / 2)
        self.eigs[0, 0] = 1e-10
        self.rank = (-self.eigs).flatten().argsort()

    def _modes(self):
        return np.arange(int(self.Nmesh) ** 2)

    # ---- device plumbing
    def _release(self):
        if self._dev is not None:
            _lib.load().ces_darcy_destroy(self._dev[0])
            self._dev = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _handle(self, want_obs):
        import torch

        obs = np.asarray(getattr(self, "obs_index", []), dtype=np.int64).reshape(-1) if want_obs else np.zeros(0, np.int64)
        key = (int(self.Nmesh), float(self.alpha), float(self.tau), int(self.p), obs.tobytes())
        if self._dev is not None and self._dev[1] == key:
            return self._dev[0]
        self._release()
        if not torch.cuda.is_available():
            raise RuntimeError("ces_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        N = int(self.Nmesh)
        if not 4 <= N <= 128:
            raise NotImplementedError("the device Darcy path covers 4 <= Nmesh <= 128 (got %d): one member must fit the "
                                      "registers and shared memory of one thread-block cluster" % N)
        phiT = np.ascontiguousarray(kl_operator(N, self.alpha, self.tau, self._modes()))
        centres = (np.arange(N) + 0.5) / N
        nodes = np.arange(N) / (N - 1.0)
        S = np.ascontiguousarray(spline_operator(centres, nodes))
        S2 = np.ascontiguousarray(spline_operator(nodes, centres))
        lib = _lib.load()
        h = ctypes.c_void_p()
        self._stream = torch.cuda.current_stream()          # the handle launches on this stream from now on
        _lib.check(lib.ces_darcy_create(N, int(self.p), _lib.host_ptr(phiT), _lib.host_ptr(S), _lib.host_ptr(S2),
                                        obs.ctypes.data if obs.size else None, int(obs.size),
                                        ctypes.c_void_p(self._stream.cuda_stream), ctypes.byref(h)))
        self._dev = (h, key)
        return h

    def _forward(self, h, U_dev, G_dev, full):
        import torch
        from .engine import stream_guard

        iters = ctypes.c_int()
        with stream_guard(torch, self._stream):
            _lib.check(_lib.load().ces_darcy_forward(h, ctypes.c_void_p(U_dev.data_ptr()), int(U_dev.stride(0)),
                                                     int(U_dev.shape[1]), ctypes.c_void_p(G_dev.data_ptr()),
                                                     int(G_dev.stride(0)), 1 if full else 0, float(self.tol),
                                                     int(self.max_iter), ctypes.byref(iters)))
        self.last_iterations = iters.value

    def evaluate_ensemble(self, engine, U_dev, G_dev):
        """G[:, j] = observations of member j (``enka.G_ens``); U_dev (p, cols), G_dev (n_obs, cols) on the GPU."""
        self._forward(self._handle(True), U_dev, G_dev, False)
        return G_dev

    def last_stats(self):
        """(members, CG iterations summed over the members, solver milliseconds) of the last device call."""
        if self._dev is None:
            return 0, 0, 0.0
        members, total, ms = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
        _lib.check(_lib.load().ces_darcy_last_stats(self._dev[0], ctypes.byref(members), ctypes.byref(total),
                                                    ctypes.byref(ms)))
        return members.value, total.value, ms.value

    def solve_ensemble(self, U, full_solution=True):
        """Host convenience: U (p, n) numpy -> (N^2, n) cell-centre pressures or (n_obs, n) observations."""
        import torch

        U = np.ascontiguousarray(np.asarray(U, dtype=np.float64).reshape(int(self.p), -1))
        n = U.shape[1]
        rows = int(self.Nmesh) ** 2 if full_solution else len(self.obs_index)
        h = self._handle(not full_solution)
        Ud = torch.from_numpy(U).cuda()
        Gd = torch.empty(rows, n, dtype=torch.float64, device="cuda")
        self._forward(h, Ud, Gd, full_solution)
        return Gd.cpu().numpy()

    def __call__(self, xi, full_solution=False):
        """One member (ces/darcy.py:20-38): flattened cell-centre pressure, or its ``obs_index`` entries."""
        out = self.solve_ensemble(np.asarray(xi, dtype=np.float64).reshape(-1, 1), full_solution=full_solution)
        return out[:, 0]


class model_trunc(model):
    """Darcy flow on the leading p KL modes (by eigenvalue rank).  ces/darcy.py:100-138."""

    def __init__(self, alpha=2., tau=3., Nmesh=2. ** 4, p=10):
        super().__init__(alpha=alpha, tau=tau, Nmesh=Nmesh)
        super().set_rank()
        self.p = p

    def _modes(self):
        return self.rank[:self.p]

    def set_initial(self, seed=1):
        np.random.seed(seed)
        ustar = np.random.normal(0, 1, int(self.Nmesh * self.Nmesh))
        self.ustar = ustar[self.rank[:self.p]]
