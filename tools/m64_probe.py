#!/usr/bin/env python
"""TMEM layout of an M = 64 (cta_group::1) accumulator and the effect of a lane offset in the D address."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from presight_b200 import _lib
lib = _lib.load()
lib.ps_tc5_m64_probe.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
torch.manual_seed(0)
X = torch.randn(128, 64, device="cuda").bfloat16().float()
Y = torch.randn(128, 64, device="cuda").bfloat16().float()
ref = X.t() @ Y            # [64 (m) x 64 (n)]
full = torch.zeros(128, 64, device="cuda")
full[:64] = ref            # the M = 128 product reads 128 "columns" of X: rows 64.. come from the Y tile behind it
full[64:] = Y.t() @ Y


def where(d, target, scale):
    """for each lane of dump d: which row of `target` (x scale) it holds, or None"""
    out = []
    for lane in range(128):
        hit = None
        for m in range(target.shape[0]):
            if torch.allclose(d[lane], scale * target[m], rtol=2e-2, atol=2e-2):
                hit = m
                break
        out.append(hit)
    return out


for lane_off in (-1, 16, 32, 64):
    dump = torch.zeros(2, 128, 64, device="cuda")
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = lib.ps_tc5_m64_probe(X.data_ptr(), Y.data_ptr(), dump.data_ptr(), cyc.data_ptr(), lane_off, None)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print(f"lane_off {lane_off}: {e}")
        break
    a = where(dump[0], ref, 1.0)
    a128 = where(dump[0], full, 1.0)
    print(f"lane_off {lane_off}: after M=64 product at lane 0: lanes holding X^T Y rows: "
          f"{[(l, m) for l, m in enumerate(a) if m is not None][:70]}")
    print(f"   lanes still holding the M=128 product rows: {[(l, m) for l, m in enumerate(a128) if m is not None and a[l] is None][:70]}")
    if lane_off >= 0:
        b2 = where(dump[1], ref, 2.0)
        b1 = where(dump[1], ref, 1.0)
        print(f"   after the second product (2x) with lane offset: lanes with 2x rows {[(l, m) for l, m in enumerate(b2) if m is not None][:70]}")
        print(f"   lanes still with 1x rows {[(l, m) for l, m in enumerate(b1) if m is not None][:70]}")
    print(f"   16 MN-major N=64 MMAs: M=64 {int(cyc[0])} cycles, M=128 {int(cyc[1])} cycles")
