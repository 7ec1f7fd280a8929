"""world_size-2 (and 3) gloo test of the column-sharded orchestration (ces_b200.engine.run_sharded_phases)
on CPU: the collectives between the phases, the rank-major gather layout and the zero-padded ragged last
shard must reproduce the single-process oracle step.  The per-rank arithmetic is a numpy stand-in
(tests/_numpy_phases.py); on the GPU the same orchestration drives libces_b200.so."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    import socket

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, rule, d, k, J, q, ts=None, t_last=None, formulation="interaction"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _numpy_phases import NumpyPhases
        from ces_b200.engine import run_phases
        from oracle import eks_oracle as eo

        pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=True)
        ph = NumpyPhases(d, k, J, rank, world, pr["y"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"])
        sl = slice(ph.lo, ph.hi)
        fixed, resolve = None, None
        if ts == "constant":
            fixed, resolve = 0.02, "always"
        elif ts == "mix":
            fixed, resolve = (0.02 if t_last >= 4.0 else None), (t_last, 1.0)
        elif ts == "spectral":
            resolve = "spectral"
        phases = ph.bind(rule, pr["U0"][:, sl], pr["G"][:, sl], pr["xi"][:, sl], fixed_h=fixed)
        run_phases(phases, ph.buffer, (dist, None, rank), d, k, rule, resolve, formulation)
        ref = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], pr["mu"], pr["Sigma0"], pr["ustar"], pr["xi"],
                      time_step=ts, delta_t=0.02, t_last=t_last)
        err = np.abs(ph.out - ref["Uk"][:, sl]).max() / np.abs(ref["Uk"]).max() if ph.cols else 0.0
        herr = abs(ph.hk - ref["hk"]) / ref["hk"]
        merr = max(abs(ph.metrics[m] - ref["metrics"][m]) / abs(ref["metrics"][m]) for m in ph.metrics)
        q.put((rank, float(err), float(herr), float(merr)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rule,J,ts,t_last", [
    (2, "aldi", 37, None, None), (2, "aldi_constant", 40, None, None), (2, "eks", 33, None, None),
    (3, "aldi", 16, None, None), (2, "eki", 21, None, None),
    (2, "aldi", 29, "constant", 0.3), (2, "eks", 26, "constant", None), (2, "aldi", 31, "mix", 1.5),
    (2, "aldi", 31, "mix", 5.0), (2, "aldi", 35, "factored", None), (3, "eks", 22, "factored", None),
    (2, "aldi_constant", 27, "factored", None), (2, "aldi", 30, "spectral", None), (3, "eks", 19, "spectral", None)])
def test_sharded_orchestration_matches_single_process(world, rule, J, ts, t_last):
    formulation = "interaction"
    if ts == "factored":
        ts, formulation = None, "factored"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rule, 5, 7, J, q, ts, t_last, formulation))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    for rank, err, herr, merr in got:
        assert err < 1e-12 and herr < 1e-12 and merr < 1e-12, (rank, err, herr, merr)
