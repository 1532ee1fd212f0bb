#!/usr/bin/env python
"""Cycles per tcgen05.mma (M = 128, K = 16, bf16, operands in shared memory) vs N, operand majorness and accumulator reuse,
issued back to back by one thread: `issue` = until the last instruction is issued, `done` = until the commit arrives."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from presight_b200 import _lib
lib = _lib.load()
lib.ps_tc5_mma_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
n = 256
for M in (128, 64):
    for mn in (0, 1):
        for N in (16, 64, 128):
            lib.ps_tc5_mma_cost(N, mn, n, 1, M, out.data_ptr(), None)
            torch.cuda.synchronize()
            lib.ps_tc5_mma_cost(N, mn, n, 1, M, out.data_ptr(), None)
            torch.cuda.synchronize()
            i, d = out.tolist()
            print(f"M={M:3d} {'MN' if mn else 'K '}-major N={N:3d}: issue {i / n:6.1f} cyc/MMA, done {d / n:6.1f} cyc/MMA")
