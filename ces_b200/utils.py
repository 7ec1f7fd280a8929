"""Map-type forward models with the constructor / call signatures of ``ces/utils.py``,
evaluated for a whole ensemble in one device call.

Reference classes mirrored (same names, attributes and single-particle ``__call__``):
``lineal`` (ces/utils.py:5-31), ``lineal_log`` (:33-51), ``elliptic`` (:53-89),
``banana`` (:91-122).  The reference evaluates them one Python call per particle
inside ``enka.G_ens`` (ces/calibrate.py:123-130); here ``sampling`` recognises the
``device_kind`` attribute and runs ``ces_forward_map`` on the ensemble resident in
HBM.  Observation noise inside the model (``flag_noise=True``) draws from the
global numpy RNG particle by particle in the reference and therefore stays on
the reference's per-particle host protocol (see ``sampling.G_ens``).
"""
import numpy as np

from .engine import Engine


class _DeviceMap(object):
    type = "map"
    device_kind = None
    n_in = None

    def _device_args(self, torch):
        """(A_dev, lda, b_dev, params) for ces_forward_map."""
        return None, 0, None, None

    def __getstate__(self):
        # device handles and buffers do not travel (joblib pickles the model when enka.parallel is set); they are rebuilt
        state = dict(self.__dict__)
        state.pop("_single_cache", None)
        if "_dev" in state:
            state["_dev"] = None
        return state

    def evaluate_ensemble(self, engine, U_dev, G_dev):
        A, lda, b, params = self._device_args(engine.torch)
        return engine.forward_map(self.device_kind, U_dev, G_dev, A=A, lda=lda, b=b, params=params)

    def _single(self, theta, n_obs):
        """One particle through the same device kernel the ensemble uses."""
        import torch

        theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        p = theta.shape[0]
        # a handle needs at least two particles: evaluate a 2-column ensemble and keep column 0.  The forward-only handle
        # and its two device buffers are kept on the model: a caller that loops over particles (ces/sample.py:121-196
        # evaluates the model once per MCMC proposal) pays one small upload, one launch and one download per call, not a
        # handle + cudaMalloc
        key = (p, int(n_obs), torch.cuda.current_device())
        cache = self.__dict__.get("_single_cache")
        if cache is None or cache[0] != key:
            if cache is not None:
                cache[1].close()
            cache = (key, Engine(p, n_obs, 2, d_panel_bytes=-1),       # forward maps only
                     torch.empty(p, 2, dtype=torch.float64, device="cuda"),
                     torch.empty(n_obs, 2, dtype=torch.float64, device="cuda"))
            self.__dict__["_single_cache"] = cache
        _, eng, U, G = cache
        U.copy_(torch.from_numpy(np.stack([theta, theta], axis=1)))
        self.evaluate_ensemble(eng, U, G)
        return G[:, 0].cpu().numpy()


class lineal(_DeviceMap):
    """G(theta) = A theta + b.  ces/utils.py:5-31."""
    device_kind = "lineal"

    def __init__(self, A, b=0, flag_noise=False):
        self.A = np.asarray(A, dtype=np.float64)
        self.b = b
        self.n_obs = self.A.shape[0]
        self.flag_noise = flag_noise
        self.noise_sigma = np.sqrt(0.1)
        self.model_name = "lineal"
        self.type = "map"
        self._dev = None

    def __repr__(self):
        return self.model_name

    def __str__(self):
        return self.model_name

    def _device_args(self, torch):
        if self._dev is None:
            k, p = self.A.shape
            ld = (p + 15) // 16 * 16
            A = torch.zeros(k, ld, dtype=torch.float64, device="cuda")
            A[:, :p] = torch.from_numpy(self.A)
            b = None
            if np.ndim(self.b) > 0 or self.b != 0:
                b = torch.from_numpy(np.broadcast_to(np.asarray(self.b, dtype=np.float64), (k,)).copy()).cuda()
            self._dev = (A, ld, b)
        A, ld, b = self._dev
        return A, ld, b, None

    def __call__(self, theta):
        out = self._single(theta, self.n_obs)
        if self.flag_noise:
            out = out + self.noise_sigma * np.random.normal()
        return out


class lineal_log(lineal):
    """G(phi) = A exp(phi).  ces/utils.py:33-51."""
    device_kind = "lineal_log"

    def __init__(self, A, flag_noise=False):
        super().__init__(A, flag_noise=flag_noise)
        self.model_name = "lineal_log"
        self.jacobian_adjusted = True

    def grad_logjacobian(self, params):
        return -np.exp(-params)

    def logjacobian(self, params):
        return -params.sum(axis=0) if self.jacobian_adjusted else 0.0


class elliptic(_DeviceMap):
    """p(x) = u2 x + exp(-u1)(x - x^2)/2 observed at x1 = 1/4, x2 = 3/4.  ces/utils.py:53-89."""
    device_kind = "elliptic"

    def __init__(self, flag_noise=False):
        self.x1 = 1. / 4
        self.x2 = 3. / 4
        self.flag_noise = flag_noise
        self.sigma = np.sqrt(0.01)
        self.model_name = "elliptic"
        self.type = "map"
        self.n_obs = 2

    def __repr__(self):
        return self.model_name

    __str__ = __repr__

    def _device_args(self, torch):
        return None, 0, None, [self.x1, self.x2]

    def __call__(self, theta, dG=False):
        if dG:
            u1, u2 = theta
            e = np.exp(-u1)
            return np.array([[-e * (-self.x1 ** 2 + self.x1) * 0.5, self.x1],
                             [-e * (-self.x2 ** 2 + self.x2) * 0.5, self.x2]])
        x, y = self._single(theta, 2)
        if self.flag_noise:
            x = x + self.sigma * np.random.normal()
            y = y + self.sigma * np.random.normal()
        return [x, y]


class banana(_DeviceMap):
    """(a u1, u2/a - b (u1^2 + a^2)).  ces/utils.py:91-122.

    The reference draws ``np.random.normal(0, 1, [2])`` in EVERY evaluation -- ``flag_noise * chol(Gamma).dot(normal)``
    (:122) multiplies the draw by zero instead of skipping it -- so a seeded script's random stream advances by two
    normals per particle per forward pass.  ``rng_draws_per_call`` tells the batched callers (``sampling.run``,
    ``enka.G_ens``, ``MCMC.model_mh``) to consume the same variates, so the noise that follows is the reference's."""
    device_kind = "banana"
    rng_draws_per_call = 2

    def __init__(self, a=1.0, b=.5, rho=.9, flag_noise=False):
        self.flag_noise = flag_noise
        self.sigma = np.sqrt(0.55)
        self.model_name = "banana"
        self.type = "map"
        self.a = a
        self.b = b
        self.n_obs = 2
        self.Gamma = (0.55 ** 2) * np.array([[1.0, rho], [rho, 1.0]])

    def __repr__(self):
        return self.model_name

    __str__ = __repr__

    def _device_args(self, torch):
        return None, 0, None, [self.a, self.b]

    def __call__(self, theta, dG=False):
        out = self._single(theta, 2)
        z = np.random.normal(0, 1, [2, ])                   # drawn unconditionally, like the reference (:122)
        if self.flag_noise:
            out = out + np.linalg.cholesky(self.Gamma).dot(z)
        return out


# ------------------------------------------------------------------------------------------------
# 'pde'-type models (ces/utils.py:124-447): Lorenz 63 and the two-scale Lorenz 96 family.
class _DeviceODE(object):
    """Common device plumbing of the Lorenz models.

    The reference integrates one particle per ``enka.G_pde`` call with scipy's adaptive integrators
    (ces/calibrate.py:132-154; ces/utils.py:168-179, 316-330).  Here ``sampling.run`` hands the whole ensemble to
    ``evaluate_ensemble_pde`` and the device integrates every particle with ``self.substeps`` classical RK4 steps per
    output interval of the uniform grid ``t`` (libces_b200: ``ces_lorenz63_forward`` / ``ces_lorenz96_forward``).  The
    systems are chaotic, so beyond a few Lyapunov times the trajectories of any two integrators differ and only the
    window statistics are comparable; ``substeps`` trades cost for accuracy.
    """
    type = 'pde'
    device_kind = None
    substeps = 10

    @staticmethod
    def _grid(t):
        t = np.asarray(t, dtype=np.float64).reshape(-1)
        if t.shape[0] < 2:
            raise ValueError("the time vector needs at least two samples")
        dt = (t[-1] - t[0]) / (t.shape[0] - 1)
        if not np.allclose(np.diff(t), dt, rtol=1e-9, atol=1e-12):
            raise ValueError("the device integrators need a uniform time vector (np.arange / np.linspace)")
        return int(t.shape[0]), float(dt)

    def _launch(self, U_dev, W0_dev, t, G_dev, Wend_dev, traj_dev):
        raise NotImplementedError

    def evaluate_ensemble_pde(self, engine, U_dev, W0_dev, t, G_dev, Wend_dev):
        """Statistics G (n_obs, cols) and final states Wend (n_state, cols) of every particle; all CUDA float64."""
        self._launch(U_dev, W0_dev, t, G_dev, Wend_dev, None)
        return G_dev, Wend_dev

    def solve(self, w0, t, args=()):
        """Trajectory (len(t), n_state) of one particle, like the reference's ``solve`` (ces/utils.py:168-179)."""
        import torch

        n_out, _ = self._grid(t)
        p = len(args)
        U = torch.tensor(np.asarray(args, dtype=np.float64).reshape(p, 1) if p else np.zeros((1, 1)), device="cuda")
        W0 = torch.tensor(np.asarray(w0, dtype=np.float64).reshape(self.n_state, 1), device="cuda")
        G = torch.empty(self._n_stats(), 1, dtype=torch.float64, device="cuda")
        traj = torch.empty(n_out * self.n_state, 1, dtype=torch.float64, device="cuda")
        self._launch(U, W0, t, G, None, traj, p=p, lenient=True)
        return traj.cpu().numpy().reshape(n_out, self.n_state)


class lorenz63(_DeviceODE):
    """Lorenz 63 with parameters (r, b), sigma = 10.  ces/utils.py:124-194."""
    device_kind = "lorenz63"
    _log_params = 0

    def __init__(self, l_window=10, freq=100):
        self.n_state = 3
        self.n_obs = 9
        self.l_window = l_window
        self.freq = freq
        self.solve_init = False
        self.model_name = 'lorenz63'
        self.type = 'pde'

    def __repr__(self):
        return self.model_name

    def __str__(self):
        return self.model_name + str(self.n_state)

    def __call__(self, w, t, r=28., b=8. / 3):
        return self.model(w, t, 10., r, b)

    def model(self, w, t, sigma=10., r=28., b=8. / 3):
        x, y, z = w
        return [sigma * (y - x), r * x - y - x * z, x * y - b * z]

    def _n_stats(self):
        return 9

    def _launch(self, U_dev, W0_dev, t, G_dev, Wend_dev, traj_dev, p=None, lenient=False):
        import ctypes
        import torch
        from . import _lib

        n_out, dt = self._grid(t)
        window = int(self.l_window * self.freq)
        if lenient and (n_out - 1) % window != 0:
            window = n_out - 1               # solve() only wants the trajectory
        p = U_dev.shape[0] if p is None else p
        vp = ctypes.c_void_p
        _lib.check(_lib.load().ces_lorenz63_forward(
            vp(torch.cuda.current_stream().cuda_stream), self._log_params, vp(U_dev.data_ptr()), int(U_dev.stride(0)), int(p),
            int(W0_dev.shape[1]), vp(W0_dev.data_ptr()), int(W0_dev.stride(0)), n_out, dt, int(self.substeps), window,
            vp(G_dev.data_ptr()), int(G_dev.stride(0)),
            vp(Wend_dev.data_ptr()) if Wend_dev is not None else None, int(Wend_dev.stride(0)) if Wend_dev is not None else 0,
            vp(traj_dev.data_ptr()) if traj_dev is not None else None, int(traj_dev.stride(0)) if traj_dev is not None else 0))

    def statistics(self, ws):
        """Window means of (x, y, z, x^2, y^2, z^2, xy, xz, yz) over the last adjacent window of t[1:]
        (ces/utils.py:181-194) of a trajectory the caller holds on the host."""
        ws = np.asarray(ws)
        xs, ys, zs = ws[:, 0], ws[:, 1], ws[:, 2]
        m = np.asarray([xs, ys, zs, xs ** 2, ys ** 2, zs ** 2, xs * ys, xs * zs, ys * zs])
        return m[:, 1:].reshape(self.n_obs, -1, int(self.l_window * self.freq)).mean(axis=2)[:, -1]


class lorenz63_log(lorenz63):
    """Lorenz 63 in (log r, log b).  ces/utils.py:196-229."""
    _log_params = 1

    def __init__(self, l_window=10, freq=100):
        super().__init__(l_window=l_window, freq=freq)
        self.model_name = 'lorenz63_log'

    def __call__(self, w, t, log_r=np.log(28.), log_b=np.log(8. / 3)):
        return self.model(w, t, 10., log_r, log_b)

    def model(self, w, t, sigma=10., log_r=np.log(28.), log_b=np.log(8. / 3)):
        return lorenz63.model(self, w, t, sigma, np.exp(log_r), np.exp(log_b))

    def grad_logjacobian(self, params):
        return -np.exp(-params)

    def logjacobian(self, params):
        return -params.sum(axis=0)


class lorenz96(_DeviceODE):
    """Two-scale Lorenz 96 with parameters (h, F, log c, b).  ces/utils.py:231-342."""
    device_kind = "lorenz96"
    substeps = 50                    # the fast variables evolve at rates ~ c b |Y| = O(100)
    _slots = (0, 1, 2, 3)            # U row i -> (h, F, log c, b)
    _out_mode, _out_col = 0, 0

    def __init__(self, n_slow=36, n_fast=10, l_window=10, freq=10, spinup=10):
        self.n_slow = n_slow
        self.n_fast = n_fast
        self.n_state = self.n_slow * (self.n_fast + 1)
        self.l_window = l_window
        self.freq = freq
        self.spinup = spinup
        self.solve_init = False
        self.model_name = 'lorenz96'
        self.type = 'pde'

    def __repr__(self):
        return self.model_name + ',' + str(self.n_slow) + ',' + str(self.n_fast)

    def __str__(self):
        print('Model: ..................... Lorenz 96')
        print('Number of slow variables ... %s' % (self.n_slow))
        print('Number of fast variables ... %s' % (self.n_fast))
        print('Number of parameters........ %s' % (len(self._slots)))
        print('Solver initialized ......... %s' % (self.solve_init))
        return str()

    @property
    def n_obs(self):
        return self._n_stats()

    def _n_stats(self):
        return 5 * self.n_slow if self._out_mode == 0 else 5

    def __call__(self, t, w, *args):
        full = [1., 10., np.log(10.), 10.]
        for slot, val in zip(self._slots, args):
            full[slot] = val
        return self.model(w, t, *full)

    def generate_initial(self):
        """ces/utils.py:277-287: slow variables uniform in [-5, 10), every fast variable equal to its slow one."""
        x0 = np.empty(self.n_slow + self.n_slow * self.n_fast)
        x0[:self.n_slow] = np.random.rand(self.n_slow) * 15 - 5
        for k in range(0, self.n_slow):
            x0[self.n_slow + k * self.n_fast: self.n_slow + (k + 1) * self.n_fast] = x0[k]
        return x0

    def model(self, X, t, h=1., F=10., log_c=np.log(10.), b=10.):
        """Right-hand side for one state vector on the host (ces/utils.py:289-308); the ensemble path never calls it."""
        c = np.exp(log_c)
        ns, nf = self.n_slow, self.n_fast
        X = np.asarray(X, dtype=np.float64)
        Y, X = X[ns:], X[:ns]
        n = ns * nf
        k = np.arange(ns)
        dX = -X[k - 1] * (X[k - 2] - X[(k + 1) % ns]) - X + F - (h * c) * Y.reshape(ns, nf).mean(axis=1)
        j = np.arange(n)
        dY = -c * b * Y[(j + 1) % n] * (Y[(j + 2) % n] - Y[j - 1]) - c * Y + ((h * c) / nf) * X[j // nf]
        return np.hstack((dX, dY))

    def set_solver(self, method='RK45', T=20, dt=0.1):
        """Kept for call compatibility (ces/utils.py:310-314); the device scheme is fixed-step RK4 (``substeps``)."""
        self.method = method
        self.dt = dt
        self.T = T
        self.solve_init = True

    def _launch(self, U_dev, W0_dev, t, G_dev, Wend_dev, traj_dev, p=None, lenient=False):
        import ctypes
        import torch
        from . import _lib

        n_out, dt = self._grid(t)
        window = int(self.l_window * self.freq)
        skip = int(self.spinup * self.freq + 1)
        if lenient and (skip >= n_out or (n_out - skip) % window != 0):
            skip, window = 1, n_out - 1
        p = U_dev.shape[0] if p is None else p
        slots = (ctypes.c_int * 4)(*(list(self._slots) + [-1] * (4 - len(self._slots))))
        vp = ctypes.c_void_p
        _lib.check(_lib.load().ces_lorenz96_forward(
            vp(torch.cuda.current_stream().cuda_stream), slots, int(self.n_slow), int(self.n_fast), vp(U_dev.data_ptr()),
            int(U_dev.stride(0)), int(p), int(W0_dev.shape[1]), vp(W0_dev.data_ptr()), int(W0_dev.stride(0)), n_out, dt,
            int(self.substeps), skip, window, int(self._out_mode), int(self._out_col), vp(G_dev.data_ptr()),
            int(G_dev.stride(0)),
            vp(Wend_dev.data_ptr()) if Wend_dev is not None else None, int(Wend_dev.stride(0)) if Wend_dev is not None else 0,
            vp(traj_dev.data_ptr()) if traj_dev is not None else None, int(traj_dev.stride(0)) if traj_dev is not None else 0))

    def _phi(self, ws):
        ws = np.asarray(ws).T
        ns, nf, W = self.n_slow, self.n_fast, int(self.l_window * self.freq)
        data = np.copy(ws[:, int(self.spinup * self.freq + 1):].reshape(self.n_state, -1, W))
        fast = data[ns:].reshape(ns, nf, -1, W)
        return np.vstack([data[:ns].mean(axis=2), (data[:ns] ** 2).mean(axis=2), fast.mean(axis=1).mean(axis=2),
                          (data[ns:] ** 2).reshape(ns, nf, -1, W).mean(axis=1).mean(axis=2),
                          (data[:ns] * fast.mean(axis=1)).mean(axis=2)])

    def statistics(self, ws):
        """ces/utils.py:332-342 for a trajectory the caller holds on the host."""
        return self._phi(ws)[:, -1]

    def grad_logjacobian(self, params):
        gradlogjac = np.zeros_like(params)
        gradlogjac[2] = -np.exp(-gradlogjac[2])
        return gradlogjac


class lorenz96_hom(lorenz96):
    """Statistics averaged over the slow index (``hom``) or taken at k = 7.  ces/utils.py:349-368."""

    def __init__(self):
        super().__init__()
        self.hom = True

    @property
    def _out_mode(self):
        return 1 if self.hom else 2

    @property
    def _out_col(self):
        return 0 if self.hom else 7

    def statistics(self, ws):
        phi = self._phi(ws)[:, -1].reshape(5, -1)
        return phi.mean(axis=1) if self.hom else phi[:, 7]


class lorenz96Fc(lorenz96):
    """Parameters (F, log c).  ces/utils.py:370-390."""
    _slots = (1, 2)

    def __init__(self):
        super().__init__()

    def __repr__(self):
        return self.model_name + ',' + str(self.n_slow) + ',' + str(self.n_fast) + ',' + str(2)


class lorenz96Fb(lorenz96):
    """Parameters (F, b).  ces/utils.py:392-409."""
    _slots = (1, 3)

    def __repr__(self):
        return self.model_name + ',' + str(self.n_slow) + ',' + str(self.n_fast) + ',' + str(2)


class lorenz96hFb(lorenz96):
    """Parameters (h, F, b).  ces/utils.py:411-428."""
    _slots = (0, 1, 3)

    def __repr__(self):
        return self.model_name + ',' + str(self.n_slow) + ',' + str(self.n_fast) + ',' + str(3)


class lorenz96hcb(lorenz96):
    """Parameters (h, log c, b).  ces/utils.py:430-447."""
    _slots = (0, 2, 3)

    def __repr__(self):
        return self.model_name + ',' + str(self.n_slow) + ',' + str(self.n_fast) + ',' + str(3)
