"""Column-sharded update over NCCL (one process per GPU) against the single-process oracle.
Needs >= 2 GPUs; skipped on a 1-GPU box (the orchestration itself is covered on CPU by
tests/test_multirank_gloo.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    import socket

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, rule, d, k, J, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from ces_b200 import calibrate, utils as cutils
        from ces_b200.engine import Engine
        from oracle import eks_oracle as eo, forward_oracle as fo

        pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=True)
        eng = Engine(d, k, J, group=dist.group.WORLD)
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        sl = slice(eng.col_lo, eng.col_hi)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, sl])).cuda()
        out, hk, met = eng.step(rule, dev(pr["U0"]), dev(pr["G"]), dev(pr["xi"]))
        ref = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"])
        err = float(np.abs(out.cpu().numpy() - ref["Uk"][:, sl]).max() / np.abs(ref["Uk"]).max()) if eng.cols else 0.0
        herr = abs(hk - ref["hk"]) / ref["hk"]
        merr = max(abs(met[m] - ref["metrics"][m]) / abs(ref["metrics"][m]) for m in met)
        eng.close()
        # the reference-facing run() loop, sharded: every rank passes the same arguments
        A = pr["A"]
        s = calibrate.sampling(d, k, J)
        s.mu, s.sigma, s.ustar, s.T, s.group = pr["mu"], pr["Sigma0"], pr["ustar"], 3, dist.group.WORLD
        np.random.seed(5)
        s.run(pr["y"], pr["U0"], cutils.lineal(A), pr["Gamma"], None, t_tol=1e9)
        np.random.seed(5)
        U, t = pr["U0"], None
        for _ in range(3):
            o = eo.step("aldi", pr["y"], U, fo.lineal(A, U), pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"],
                        np.random.normal(0, 1, [d, J]), t_last=t)
            U, t = o["Uk"], o["t"]
        rerr = float(np.abs(s.Ustar - U).max() / np.abs(U).max())
        q.put((rank, err, float(herr), float(merr), rerr, s.Uall.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rule,J", [(2, "aldi", 301), (2, "aldi_constant", 256), (2, "eks", 130), (4, "aldi", 1030),
                                          (4, "eks", 515), (3, "aldi_constant", 700)])
def test_multi_gpu_step_matches_oracle(world, rule, J):
    """world >= 3 exercises the batched launches over the other ranks' source blocks (rotated order with wrap-around)."""
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rule, 24, 40, J, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for rank, err, herr, merr, rerr, shape in sorted(q.get(timeout=10) for _ in range(world)):
        assert err < 1e-10 and herr < 1e-10 and merr < 1e-10, (rank, err, herr, merr)
        assert rerr < 1e-9 and shape == (4, 24, J)
