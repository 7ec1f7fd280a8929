"""The C-ABI library loads and exports every symbol include/ces_b200.h declares (CPU; no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ces_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ces_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from ces_b200 import _lib

    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from ces_b200 import _lib

    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ces_version()


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    from ces_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.ces_create(0, 5, 10, 10, 0, 1, 10, None, 0, ctypes.byref(h)) == _lib.CES_ERR_INVALID
    assert lib.ces_create(2, 5, 10, 10, 3, 2, 10, None, 0, ctypes.byref(h)) == _lib.CES_ERR_INVALID
    assert lib.ces_create(2, 5, 4, 10, 0, 2, 4, None, 0, ctypes.byref(h)) == _lib.CES_ERR_INVALID   # 2*4 < 10
    assert b"inconsistent" in lib.ces_last_error()
    assert lib.ces_phase3_interact(None, 1, 0) == _lib.CES_ERR_INVALID
    assert lib.ces_gemm(None, 7, 0, 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4) == _lib.CES_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.CES_ERR_INVALID)


def test_status_mapping_follows_numpy_conventions():
    import numpy as np
    from ces_b200 import _lib

    with pytest.raises(np.linalg.LinAlgError):
        _lib.check(_lib.CES_ERR_NOT_SPD)
    with pytest.raises(MemoryError):
        _lib.check(_lib.CES_ERR_NOMEM)
    with pytest.raises(_lib.CesError):
        _lib.check(_lib.CES_ERR_CUDA)


def test_sass_uses_fp64_tensor_cores_and_tma():
    """cuobjdump evidence that the GEMM is DMMA + TMA (skipped if cuobjdump is unavailable)."""
    import shutil
    import subprocess
    from ces_b200 import _lib

    exe = shutil.which("cuobjdump")
    if exe is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "DMMA.8x8x4" in sass
    assert "UTMALDG.2D" in sass
    assert "arch = sm_100a" in sass


def test_host_chunk_schedule_is_rank_independent_and_well_formed():
    """ces_host_chunk_schedule decides how the G upload of a host step is cut into row chunks.  Every rank of a column-sharded
    step all-reduces the means chunk by chunk, so all ranks must compute the SAME chunks: the function takes no rank, no
    column count and (for nranks > 1) no measured rate -- a rank-local measurement made two ranks disagree on the chunk
    count once and the step hung in its all-reduces.  Host-only code: runs without a GPU."""
    from ces_b200 import _lib

    lib = _lib.load()

    def sched(k, Jl, panel, nranks, gbs=0.0):
        b = (ctypes.c_int64 * 9)()
        n = lib.ces_host_chunk_schedule(k, Jl, panel, nranks, gbs, b)
        return n, list(b)

    for (k, Jl, panel, nranks) in [(4096, 65536, 16384, 1), (4096, 32768, 16384, 2), (4096, 16384, 16384, 4),
                                   (4096, 8192, 8192, 8), (4096, 2048, 2048, 8), (16384, 8192, 8192, 8), (272, 2150, 1536, 2),
                                   (300, 5000, 5000, 3), (255, 65536, 16384, 1), (4096, 2047, 2047, 4)]:
        n, b = sched(k, Jl, panel, nranks)
        assert 1 <= n <= 8 and b[0] == 0 and b[n] == k and all(b[i] == k for i in range(n, 9)), (k, Jl, nranks, n, b)
        assert all(b[i] < b[i + 1] for i in range(n)) and all(b[i] % 16 == 0 for i in range(n)), b
        if k < 256 or Jl < 2048:
            assert n == 1
        assert sched(k, Jl, panel, nranks) == (n, b)                  # deterministic
    # one GPU, fast uploads: three geometric chunks; eight ranks, slow uploads: an even split
    assert sched(4096, 65536, 16384, 1)[0] == 3 and sched(4096, 8192, 8192, 8)[0] == 8
    # a measured rate only matters on a single GPU path; the nominal one is used otherwise by the caller
    assert sched(4096, 65536, 16384, 1, 55.0)[0] == 3 and sched(4096, 65536, 16384, 1, 5.0)[0] > 3


def test_header_is_plain_c():
    """include/ces_b200.h is the drop-in boundary for hosts in any language: it must parse as C99 (no C++ types, no torch
    types in the signatures) and as C++."""
    import shutil
    import subprocess
    import tempfile

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "h.c")
        with open(src, "w") as fh:
            fh.write('#include "ces_b200.h"\nint main(void) { return (int)sizeof(ces_handle_t) * 0; }\n')
        inc = os.path.join(ROOT, "include")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", src])
        if shutil.which("g++"):
            subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", src])
