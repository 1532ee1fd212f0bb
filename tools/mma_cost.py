#!/usr/bin/env python
"""Cycles per tcgen05.mma (K = 16, bf16, operands in shared memory, chunk-major / no swizzle) vs N, operand majorness,
number of independent accumulators (round-robin: 1 = every instruction depends on the previous one's accumulator) and
number of issuing threads: `issue` = until issuer 0 has issued its last instruction, `done` = until all commits arrive."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from presight_b200 import _lib
lib = _lib.load()
lib.ps_tc5_mma_cost.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
n = 256


def run(N, mn, n_acc, M, iss):
    for _ in range(2):
        rc = lib.ps_tc5_mma_cost(N, mn, n, n_acc, M, iss, out.data_ptr(), None)
        assert rc == 0, _lib.last_error() if hasattr(_lib, "last_error") else rc
        torch.cuda.synchronize()
    i, d = out.tolist()
    tot = n * iss
    print(f"M={M:3d} {'MN' if mn else 'K '}-major N={N:3d} acc={n_acc} issuers={iss}: issue {i / n:6.1f} cyc/MMA/issuer, "
          f"done {d / tot:6.1f} cyc/MMA", flush=True)


for M in (128, 64):
    for mn in (0, 1):
        for N in (16, 64, 128, 256):
            for n_acc in (1, 2, 4):
                if n_acc * N > 512 or (mn and N > 128):
                    continue
                run(N, mn, n_acc, M, 1)
for mn in (0, 1):
    for N in (64, 128):
        for iss in (2, 4):
            if iss * N <= 512:
                run(N, mn, iss, 128, iss)
            run(N, mn, 1, 128, iss)
