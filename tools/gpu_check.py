"""Development check run on a B200 through gpurun: kernel-level correctness against torch fp64,
step parity against the numpy oracle, and first timings (DMMA GEMM vs the cuBLAS DGEMM yardstick).
Writes gpurun_out/check.json.  Not part of the product; the judged tests live in tests/.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200 import _lib  # noqa: E402
from ces_b200.engine import Engine  # noqa: E402
from oracle import eks_oracle as eo  # noqa: E402

OUT = {}
dev = torch.device("cuda:0")
lib = _lib.load()
print(lib.ces_version().decode(), torch.cuda.get_device_name(0), flush=True)


def pad_ld(c):
    return (max(c, 1) + 15) // 16 * 16


def dmat(rows, cols, gen):
    """rows x cols random matrix in a padded buffer (ld multiple of 16); returns (buffer, view)."""
    buf = torch.zeros(rows, pad_ld(cols), dtype=torch.float64, device=dev)
    buf[:, :cols] = torch.randn(rows, cols, dtype=torch.float64, device=dev, generator=gen)
    return buf, buf[:, :cols]


def gemm(a_mode, b_mode, M, N, K, A, B, C, alpha=1.0, beta=0.0):
    st = torch.cuda.current_stream().cuda_stream
    s = lib.ces_gemm(ctypes.c_void_p(st), a_mode, b_mode, M, N, K, alpha, ctypes.c_void_p(A.data_ptr()), A.stride(0),
                     ctypes.c_void_p(B.data_ptr()), B.stride(0), beta, ctypes.c_void_p(C.data_ptr()), C.stride(0))
    _lib.check(s)


def check_gemm():
    gen = torch.Generator(device=dev).manual_seed(0)
    res = []
    shapes = [(128, 128, 16), (128, 128, 64), (128, 256, 100), (100, 70, 33), (257, 130, 1000), (2, 100, 10),
              (64, 1, 7), (1000, 1000, 50), (300, 513, 2050)]
    for (M, N, K) in shapes:
        for am in (0, 1):
            for bm in (0, 1):
                Ab, A = dmat(M, K, gen) if am == 0 else dmat(K, M, gen)
                Bb, B = dmat(K, N, gen) if bm == 0 else dmat(N, K, gen)
                Cb, C = dmat(M, N, gen)
                C0 = C.clone()
                opA = A if am == 0 else A.t()
                opB = B if bm == 0 else B.t()
                ref = 0.7 * (opA @ opB) + 0.3 * C0
                gemm(am, bm, M, N, K, Ab, Bb, Cb, 0.7, 0.3)
                torch.cuda.synchronize()
                err = float((C - ref).abs().max() / ref.abs().max())
                pad_ok = bool((Cb[:, N:] == 0).all())
                res.append(dict(M=M, N=N, K=K, a_mode=am, b_mode=bm, rel_err=err, padding_untouched=pad_ok))
                print("gemm", M, N, K, am, bm, "%.2e" % err, pad_ok, flush=True)
    OUT["gemm_correctness"] = res
    OUT["gemm_max_rel_err"] = max(r["rel_err"] for r in res)


def time_fn(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    return min(ts), float(np.median(ts))


def bench_gemm():
    gen = torch.Generator(device=dev).manual_seed(1)
    res = []
    for (M, N, K, am, bm, label) in [(8192, 8192, 4096, 1, 0, "D=E^T W"), (1024, 8192, 8192, 0, 0, "V=Ut D"),
                                     (8192, 8192, 4096, 0, 0, "NN"), (4096, 4096, 16384, 0, 1, "NT"),
                                     (16384, 16384, 4096, 1, 0, "D=E^T W cfg3")]:
        A = torch.randn((M, K) if am == 0 else (K, M), dtype=torch.float64, device=dev, generator=gen)
        B = torch.randn((K, N) if bm == 0 else (N, K), dtype=torch.float64, device=dev, generator=gen)
        C = torch.empty(M, N, dtype=torch.float64, device=dev)
        opA = A if am == 0 else A.t()
        opB = B if bm == 0 else B.t()
        flops = 2.0 * M * N * K
        t_ours = time_fn(lambda: gemm(am, bm, M, N, K, A, B, C))
        t_cublas = time_fn(lambda: torch.matmul(opA, opB, out=C))
        r = dict(label=label, M=M, N=N, K=K, ours_ms_best=t_ours[0], ours_ms_median=t_ours[1],
                 cublas_ms_best=t_cublas[0], cublas_ms_median=t_cublas[1],
                 ours_tflops=flops / t_ours[0] * 1e-9, cublas_tflops=flops / t_cublas[0] * 1e-9)
        res.append(r)
        print("bench", r, flush=True)
        del A, B, C
    OUT["gemm_bench"] = res


def check_chol():
    gen = torch.Generator(device=dev).manual_seed(2)
    res = []
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n in (3, 40, 64, 65, 200, 1024, 2048):
        Q = torch.randn(n, n + 5, dtype=torch.float64, device=dev, generator=gen)
        A = Q @ Q.t() / n + 0.1 * torch.eye(n, dtype=torch.float64, device=dev)
        buf = torch.zeros(n, pad_ld(n), dtype=torch.float64, device=dev)
        buf[:, :n] = A
        _lib.check(lib.ces_potrf(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n))
        Lref = torch.linalg.cholesky(A)
        err = float((buf[:, :n] - Lref).abs().max() / Lref.abs().max())
        # posv
        nrhs = 37
        Bb = torch.zeros(n, pad_ld(nrhs), dtype=torch.float64, device=dev)
        Bb[:, :nrhs] = torch.randn(n, nrhs, dtype=torch.float64, device=dev, generator=gen)
        B0 = Bb[:, :nrhs].clone()
        buf[:, :n] = A
        _lib.check(lib.ces_posv(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n, ctypes.c_void_p(Bb.data_ptr()),
                                Bb.stride(0), nrhs))
        Xref = torch.linalg.solve(A, B0)
        err2 = float((Bb[:, :nrhs] - Xref).abs().max() / Xref.abs().max())
        res.append(dict(n=n, potrf_rel_err=err, posv_rel_err=err2))
        print("chol", n, "%.2e %.2e" % (err, err2), flush=True)

    def run():
        buf[:, :n] = A
        _lib.check(lib.ces_potrf(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n))
    for n in (1024,):
        Q = torch.randn(n, n + 5, dtype=torch.float64, device=dev, generator=gen)
        A = Q @ Q.t() / n + 0.1 * torch.eye(n, dtype=torch.float64, device=dev)
        buf = torch.zeros(n, pad_ld(n), dtype=torch.float64, device=dev)
        t = time_fn(run)
        t2 = time_fn(lambda: torch.linalg.cholesky(A))
        OUT["potrf_1024_ms"] = dict(ours=t[0], cusolver=t2[0])
        print("potrf 1024 ms", t, t2, flush=True)
    # non-SPD detection
    n = 100
    A = torch.eye(n, dtype=torch.float64, device=dev)
    A[70, 70] = -1.0
    buf = torch.zeros(n, pad_ld(n), dtype=torch.float64, device=dev)
    buf[:, :n] = A
    s = lib.ces_potrf(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n)
    res.append(dict(non_spd_status=int(s), message=lib.ces_last_error().decode()))
    print("non-spd", s, lib.ces_last_error().decode(), flush=True)
    OUT["chol"] = res


def check_step():
    res = []
    cases = [(2, 10, 100), (64, 50, 1024), (40, 30, 17), (3, 5, 33), (256, 512, 2048), (130, 257, 1000)]
    for (d, k, J) in cases:
        for dense_g in (False, True):
            for dense_s in (False, True):
                pr = eo.linear_gaussian_problem(d, k, J, dense_gamma=dense_g)
                rng = np.random.default_rng(5)
                if dense_s:
                    S = rng.standard_normal((d, d))
                    Sigma0 = 50 * np.eye(d) + S @ S.T
                    mu = rng.standard_normal((d, 1))
                else:
                    Sigma0, mu = pr["Sigma0"], pr["mu"]
                eng = Engine(d, k, J)
                eng.set_problem(pr["y"], pr["Gamma"], Sigma0, mu, pr["ustar"])
                for rule in ("aldi", "eks", "aldi_constant", "eki"):
                    for fixed in (None, 0.05):
                        if rule == "aldi_constant" and fixed is not None:
                            continue
                        o = eo.step(rule, pr["y"], pr["U0"], pr["G"], pr["Gamma"], mu, Sigma0, pr["ustar"], pr["xi"],
                                    time_step=None if fixed is None else "constant", delta_t=fixed)
                        if fixed is not None and rule in ("eks", "aldi"):
                            # the reference re-solves D with hk*Cpp + Gamma for 'constant'; the device path under
                            # test here is the fixed-h update with the Gamma-only D -> compare against that
                            E, R, W, D = eo.interaction(pr["G"], pr["y"], pr["Gamma"])
                            o = _fixed_h_oracle(rule, pr, mu, Sigma0, D, fixed)
                        Uk, hk, met = eng.step_host(rule, pr["U0"], pr["G"], pr["xi"], fixed_h=fixed)
                        err = float(np.abs(Uk - o["Uk"]).max() / np.abs(o["Uk"]).max())
                        herr = abs(hk - o["hk"]) / abs(o["hk"])
                        merr = max(abs(met[q] - o["metrics"][q]) / abs(o["metrics"][q]) for q in met)
                        res.append(dict(d=d, k=k, J=J, dense_gamma=dense_g, dense_sigma=dense_s, rule=rule, fixed_h=fixed,
                                        U_rel_err=err, hk_rel_err=herr, metrics_rel_err=merr))
                        print("step", d, k, J, dense_g, dense_s, rule, fixed, "%.2e %.2e %.2e" % (err, herr, merr), flush=True)
                eng.close()
    OUT["step_parity"] = res
    OUT["step_max_U_rel_err"] = max(r["U_rel_err"] for r in res)
    OUT["step_max_hk_rel_err"] = max(r["hk_rel_err"] for r in res)
    OUT["step_max_metrics_rel_err"] = max(r["metrics_rel_err"] for r in res)


def _fixed_h_oracle(rule, pr, mu, Sigma0, D, h):
    U, xi = pr["U0"], pr["xi"]
    p, J = U.shape
    ubar = U.mean(axis=1)[:, None]
    Ut = U - ubar
    met = eo.metrics_cheap(U, pr["ustar"], *eo.interaction(pr["G"], pr["y"], pr["Gamma"])[:2], pr["Gamma"])
    if rule == "aldi":
        C = np.cov(U).reshape(p, p) + 1e-8 * np.eye(p)
        Uk = U - h * (Ut @ D) - h * (C @ np.linalg.solve(Sigma0, U - mu)) + h * (p + 1.0) / J * Ut \
            + np.sqrt(2 * h) * (np.linalg.cholesky(C) @ xi)
    else:
        C = np.cov(U, bias=True).reshape(p, p) + 1e-8 * np.eye(p)
        lhs = np.eye(p) + h * np.linalg.solve(Sigma0.T, C.T).T
        rhs = U - h * (Ut @ D) + h * (C @ np.linalg.solve(Sigma0, mu))
        Uk = np.linalg.solve(lhs, rhs) + np.sqrt(2 * h) * (np.linalg.cholesky(C) @ xi)
    return dict(Uk=Uk, hk=h, metrics=met)


def bench_step():
    res = []
    for (d, k, J) in [(1024, 4096, 16384)]:
        pr = eo.linear_gaussian_problem(d, k, J)
        eng = Engine(d, k, J)
        eng.set_problem(pr["y"], pr["Gamma"], pr["Sigma0"], pr["mu"], pr["ustar"])
        U = torch.from_numpy(pr["U0"]).to(dev)
        G = torch.from_numpy(pr["G"]).to(dev)
        xi = torch.from_numpy(pr["xi"]).to(dev)
        out = torch.empty_like(U)
        n0 = eng.launch_count()
        t = time_fn(lambda: eng.step("aldi", U, G, xi, out=out), iters=3, warm=1)
        launches = (eng.launch_count() - n0) // 4
        flops = eo.algorithmic_flops(J, d, k)
        r = dict(d=d, k=k, J=J, ms_best=t[0], ms_median=t[1], tflops=flops / t[0] * 1e-9, launches_per_step=launches,
                 particle_updates_per_s=J / (t[0] * 1e-3))
        res.append(r)
        print("step bench", r, flush=True)
        eng.close()
    OUT["step_bench"] = res


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "chol", "step", "bench_gemm", "bench_step"]
    t0 = time.time()
    for name in which:
        try:
            {"gemm": check_gemm, "chol": check_chol, "step": check_step, "bench_gemm": bench_gemm,
             "bench_step": bench_step}[name]()
        except Exception as exc:  # keep going: one GPU call should report as much as possible
            import traceback
            traceback.print_exc()
            OUT[name + "_error"] = repr(exc)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/check_%s.json" % "_".join(which), "w") as fh:
            json.dump(OUT, fh, indent=1)
    print("done in %.1fs" % (time.time() - t0))
