"""Kernel-level checks on the B200 through the C ABI: the FP64 DMMA GEMM in its four operand layouts,
the blocked Cholesky and the SPD solve, against torch fp64 (cuBLAS / cuSOLVER) on the same device."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ces_b200 import _lib  # noqa: E402


def _pad(c):
    return (max(c, 1) + 15) // 16 * 16


def _mat(rows, cols, gen):
    buf = torch.zeros(rows, _pad(cols), dtype=torch.float64, device="cuda")
    buf[:, :cols] = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=gen)
    return buf, buf[:, :cols]


def _gemm(lib, am, bm, M, N, K, A, B, C, alpha, beta):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.ces_gemm(st, am, bm, M, N, K, alpha, ctypes.c_void_p(A.data_ptr()), A.stride(0),
                            ctypes.c_void_p(B.data_ptr()), B.stride(0), beta, ctypes.c_void_p(C.data_ptr()), C.stride(0)))


@pytest.mark.parametrize("shape", [(128, 128, 16), (100, 70, 33), (257, 130, 1000), (2, 100, 10), (64, 1, 7),
                                   (1, 1, 1), (300, 513, 2050), (1024, 384, 4096)])
@pytest.mark.parametrize("am", [0, 1])
@pytest.mark.parametrize("bm", [0, 1])
def test_dmma_gemm_all_layouts(shape, am, bm):
    lib = _lib.load()
    M, N, K = shape
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + am * 2 + bm)
    Ab, A = _mat(M, K, gen) if am == 0 else _mat(K, M, gen)
    Bb, B = _mat(K, N, gen) if bm == 0 else _mat(N, K, gen)
    Cb, C = _mat(M, N, gen)
    C0 = C.clone()
    ref = 0.7 * ((A if am == 0 else A.t()) @ (B if bm == 0 else B.t())) + 0.3 * C0
    _gemm(lib, am, bm, M, N, K, Ab, Bb, Cb, 0.7, 0.3)
    torch.cuda.synchronize()
    assert float((C - ref).abs().max() / ref.abs().max()) < 1e-13
    assert bool((Cb[:, N:] == 0).all())          # nothing written outside the N columns


def test_gemm_rejects_misaligned_operands():
    lib = _lib.load()
    A = torch.zeros(8, 7, dtype=torch.float64, device="cuda")       # odd leading dimension
    B = torch.zeros(7, 8, dtype=torch.float64, device="cuda")
    C = torch.zeros(8, 8, dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    s = lib.ces_gemm(st, 0, 0, 8, 8, 7, 1.0, ctypes.c_void_p(A.data_ptr()), 7, ctypes.c_void_p(B.data_ptr()), 8, 0.0,
                     ctypes.c_void_p(C.data_ptr()), 8)
    assert s == _lib.CES_ERR_ALIGN


@pytest.mark.parametrize("n", [1, 3, 40, 64, 65, 200, 1024])
def test_blocked_cholesky_and_solve(n):
    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(n)
    Q = torch.randn(n, n + 5, dtype=torch.float64, device="cuda", generator=gen)
    A = Q @ Q.t() / n + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    buf = torch.zeros(n, _pad(n), dtype=torch.float64, device="cuda")
    buf[:, :n] = A
    _lib.check(lib.ces_potrf(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n))
    L = torch.linalg.cholesky(A)
    assert float((buf[:, :n] - L).abs().max() / L.abs().max()) < 1e-12
    nrhs = 37
    Bb = torch.zeros(n, _pad(nrhs), dtype=torch.float64, device="cuda")
    Bb[:, :nrhs] = torch.randn(n, nrhs, dtype=torch.float64, device="cuda", generator=gen)
    B0 = Bb[:, :nrhs].clone()
    buf[:, :n] = A
    _lib.check(lib.ces_posv(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n, ctypes.c_void_p(Bb.data_ptr()),
                            Bb.stride(0), nrhs))
    X = torch.linalg.solve(A, B0)
    assert float((Bb[:, :nrhs] - X).abs().max() / X.abs().max()) < 1e-11


def test_cholesky_reports_the_failing_pivot():
    lib = _lib.load()
    n = 100
    A = torch.eye(n, dtype=torch.float64, device="cuda")
    A[70, 70] = -1.0
    buf = torch.zeros(n, _pad(n), dtype=torch.float64, device="cuda")
    buf[:, :n] = A
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    s = lib.ces_potrf(st, ctypes.c_void_p(buf.data_ptr()), buf.stride(0), n)
    assert s == _lib.CES_ERR_NOT_SPD and b"pivot 71" in lib.ces_last_error()
    with pytest.raises(np.linalg.LinAlgError):
        _lib.check(s)


def test_device_normal_noise_is_standard_normal_and_shard_invariant():
    """ces_fill_normal: moments of N(0,1), reproducible by (seed, step), different across steps, and a column shard
    draws exactly the columns a single GPU would."""
    from ces_b200.engine import Engine

    eng = Engine(8, 4, 200000)
    try:
        a = eng.normal_noise(8, seed=7, step=3)
        b = eng.normal_noise(8, seed=7, step=3)
        c = eng.normal_noise(8, seed=7, step=4)
        assert torch.equal(a, b) and not torch.equal(a, c)
        x = a.flatten()
        assert abs(float(x.mean())) < 5e-3 and abs(float(x.var()) - 1.0) < 1e-2
        assert abs(float((x ** 3).mean())) < 2e-2 and abs(float((x ** 4).mean()) - 3.0) < 5e-2
        assert abs(float((a[0] * a[1]).mean())) < 1e-2 and abs(float((a[0, :-1] * a[0, 1:]).mean())) < 1e-2
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        part = torch.empty(8, 1000, dtype=torch.float64, device="cuda")
        _lib.check(eng.lib.ces_fill_normal(st, 7, 3, ctypes.c_void_p(part.data_ptr()), part.stride(0), 8, 1000, 5000))
        assert torch.equal(part, a[:, 5000:6000])
        # odd offsets and odd widths (J = 1000 on 8 ranks gives shards of 125 columns) draw the same columns too
        for off, w in ((125, 125), (4999, 1), (4999, 2), (7, 1000)):
            part = torch.full((8, w), float("nan"), dtype=torch.float64, device="cuda")
            _lib.check(eng.lib.ces_fill_normal(st, 7, 3, ctypes.c_void_p(part.data_ptr()), part.stride(0), 8, w, off))
            assert torch.equal(part, a[:, off:off + w]), (off, w)
    finally:
        eng.close()
