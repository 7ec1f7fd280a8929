"""One D = (1/J) E^T W launch at the cfg3 shape through ces_gemm, for ncu traffic experiments
(CES_GEMM_GROUP_M selects the rasterisation group)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200 import _lib  # noqa: E402

lib = _lib.load()
J, k = 16384, 4096
gen = torch.Generator(device="cuda").manual_seed(0)
E = torch.randn(k, J, dtype=torch.float64, device="cuda", generator=gen)
W = torch.randn(k, J, dtype=torch.float64, device="cuda", generator=gen)
D = torch.empty(J, J, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    _lib.check(lib.ces_gemm(st, 1, 0, J, J, k, 1.0 / J, ctypes.c_void_p(E.data_ptr()), J, ctypes.c_void_p(W.data_ptr()), J, 0.0,
                            ctypes.c_void_p(D.data_ptr()), J))
torch.cuda.synchronize()
print("ok", float(D[0, 0]))
