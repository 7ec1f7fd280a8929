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
