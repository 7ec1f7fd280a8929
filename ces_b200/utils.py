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

    def evaluate_ensemble(self, engine, U_dev, G_dev):
        A, lda, b, params = self._device_args(engine.torch)
        return engine.forward_map(self.device_kind, U_dev, G_dev, A=A, lda=lda, b=b, params=params)

    def _single(self, theta, n_obs):
        """One particle through the same device kernel the ensemble uses."""
        import torch

        theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        p = theta.shape[0]
        # a handle needs at least two particles: evaluate a 2-column ensemble and keep column 0
        eng = Engine(p, n_obs, 2, d_panel_bytes=-1)       # forward maps only
        try:
            U = torch.from_numpy(np.stack([theta, theta], axis=1)).cuda()
            G = torch.empty(n_obs, 2, dtype=torch.float64, device="cuda")
            self.evaluate_ensemble(eng, U, G)
            return G[:, 0].cpu().numpy()
        finally:
            eng.close()


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
    """(a u1, u2/a - b (u1^2 + a^2)).  ces/utils.py:91-122."""
    device_kind = "banana"

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
        if self.flag_noise:
            out = out + np.linalg.cholesky(self.Gamma).dot(np.random.normal(0, 1, [2, ]))
        return out
