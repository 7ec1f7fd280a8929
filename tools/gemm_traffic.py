"""One D = (1/J) E^T W launch through ces_gemm for ncu captures (traffic, pipe utilisation):
    python tools/gemm_traffic.py [cfg3|target]
cfg3: 16384 x 16384 x 4096; target: one 65536 x 16384 x 4096 column panel (what a target step launches four times).
CES_GEMM_GROUP_M selects the rasterisation group."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ces_b200 import _lib  # noqa: E402

lib = _lib.load()
shape = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
M, N, k = (65536, 16384, 4096) if shape == "target" else (16384, 16384, 4096)
gen = torch.Generator(device="cuda").manual_seed(0)
E = torch.randn(k, M, dtype=torch.float64, device="cuda", generator=gen)
W = torch.randn(k, M, dtype=torch.float64, device="cuda", generator=gen)
D = torch.empty(M, N, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    _lib.check(lib.ces_gemm(st, 1, 0, M, N, k, 1.0 / M, ctypes.c_void_p(E.data_ptr()), M, ctypes.c_void_p(W.data_ptr()), M, 0.0,
                            ctypes.c_void_p(D.data_ptr()), N))
torch.cuda.synchronize()
print("ok", float(D[0, 0]))
