#!/usr/bin/env python
"""Fixed costs (cycles) of one GEMM -> epilogue phase of the fused tcgen05 kernels, single CTA (ps_tc5_phase_lat)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from presight_b200 import _lib
lib = _lib.load()
lib.ps_tc5_phase_lat.argtypes = [C.c_void_p, C.c_void_p]
out = torch.zeros(32, dtype=torch.int64, device="cuda")
for _ in range(3):
    assert lib.ps_tc5_phase_lat(out.data_ptr(), None) == 0
    torch.cuda.synchronize()
t = out.tolist()
for name, s in (("1 MMA N=64 K-major", 0), ("4 MMA N=64 K-major", 3), ("16 MMA N=64 K-major", 6), ("16 MMA N=16 K-major", 9),
                ("16 MMA N=64 MN-major", 12)):
    print(f"{name:24s}: issue {t[s + 1] - t[s]:5d}  issue->complete seen {t[s + 2] - t[s]:5d} cycles")
print(f"tcgen05.ld x32 + wait       : {t[16] - t[15]} cycles")
print(f"relu/pack + 4 STS.128       : {t[17] - t[16]} cycles")
print(f"fence.proxy.async           : {t[18] - t[17]} cycles")
print(f"tcgen05 fence + bar.sync 128: {t[19] - t[18]} cycles")
print(f"2 x tcgen05.ld x32 + wait   : {t[20] - t[19]} cycles")
