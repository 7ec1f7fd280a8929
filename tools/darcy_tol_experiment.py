import sys, numpy as np
sys.path.insert(0, '.')
from ces_b200 import darcy as cdarcy
from oracle import darcy_oracle as do
for (N, p, scale) in [(64, 64, 1.0), (64, 64, 10.0), (128, 256, 1.0), (128, 256, 10.0), (96, 64, 3.0)]:
    rng = np.random.default_rng(N + p)
    members = 3
    U = scale * rng.standard_normal((p, members))
    ref = do.ModelTrunc(Nmesh=N, p=p)
    want = np.stack([ref(U[:, j], full_solution=True) for j in range(members)], axis=1)
    for tol in (1e-13, 1e-12, 1e-11, 1e-10, 1e-9):
        m = cdarcy.model_trunc(Nmesh=N, p=p)
        m.tol = tol
        got = m.solve_ensemble(U, full_solution=True)
        err = float(np.abs(got - want).max() / np.abs(want).max())
        _, its, _ = m.last_stats()
        print('N', N, 'scale', scale, 'tol', tol, 'relerr %.2e' % err, 'mean its %.1f' % (its / members), flush=True)
