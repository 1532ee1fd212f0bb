#!/usr/bin/env python
"""Quick on-GPU check of the tcgen05 forward MLP (precision 2) against the mma.sync bf16 path and the fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from presight_b200 import ops
import oracle as O

def run(n_in, hidden, n_layers, n_out, act, P):
    g = torch.Generator().manual_seed(n_in * 131 + n_out)
    dims = [n_in] + [hidden] * (n_layers - 1) + [n_out]
    ws = [torch.randn(dims[i + 1], dims[i], generator=g) / np.sqrt(dims[i]) for i in range(n_layers)]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(n_layers)]
    x = torch.randn(P, n_in, generator=g)
    wg, bg = [w.cuda() for w in ws], [b.cuda() for b in bs]
    y1 = ops.mlp(x.cuda(), wg, bg, act, 1)
    y2 = ops.mlp(x.cuda(), wg, bg, act, 2)
    torch.cuda.synchronize()
    y16 = O.mlp_forward_bf16_emulated(x, O.Mlp(ws, bs, "sigmoid" if act == 2 else None))
    e12 = float((y1 - y2).abs().max() / y1.abs().max())
    e2o = float((y2.cpu() - y16).abs().max() / y16.abs().max())
    print(f"{n_in}->{hidden}x{n_layers-1}->{n_out} P={P}: tc5 vs mma.sync {e12:.2e}, tc5 vs bf16 emulation {e2o:.2e}", flush=True)

if __name__ == "__main__":
    for shp in [(8, 64, 2, 1, 0), (32, 64, 2, 80, 0), (64, 64, 3, 64, 0), (47, 64, 3, 3, 2), (8, 0, 1, 1, 0), (10, 16, 2, 1, 0)]:
        for P in (1000, 128 * 700 + 5):
            run(*shp, P)
