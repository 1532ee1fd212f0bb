#!/usr/bin/env python
"""Main-grid gather / scatter time as a function of the number of ray slices the batch is cut into (L2 footprint of a
level pass = slice positions + slice feature gradients + one level's table)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from presight_b200 import fused, ops, synthetic
from presight_b200._lib import call, ptr, host_floats, stream

dev = "cuda"
n, S, L, F, log2T = 65536, 64, 16, 2, 22
rays = synthetic.make_rays(n, seed=1)
o, d = rays["origins"].to(dev), rays["directions"].to(dev)
aabb = [float(v) for v in synthetic.tile_aabb().reshape(-1)]
gg = np.exp((np.log(2048) - np.log(16)) / (L - 1))
sc = [float(np.floor(16 * gg ** l)) for l in range(L)]
nears, fars = torch.full((n, 1), 0.005, device=dev), torch.full((n, 1), 50.0, device=dev)
sp, eu = ops.spaced_bins(nears, fars, S, 5.0, torch.rand(n, 1, device=dev))
x01, sel = fused._ray_points(o, d, eu.contiguous(), aabb, True)
table = (torch.rand(L << log2T, F, device=dev) * 2 - 1) * 1e-3
dtable = torch.zeros_like(table)
P = n * S
for k in (1, 2, 3, 4, 6, 8, 12, 16):
    step = (n + k - 1) // k
    feats = [torch.empty(min(step, n - c0) * S * L * F, device=dev) for c0 in range(0, n, step)]
    douts = [torch.randn_like(f) for f in feats]

    def fwd():
        for i, c0 in enumerate(range(0, n, step)):
            c1 = min(c0 + step, n)
            call("ps_hash_fwd_lm", ptr(x01[c0 * S:c1 * S]), (c1 - c0) * S, ptr(table), host_floats(sc), L, F, log2T, ptr(feats[i]), stream())

    def bwd():
        for i, c0 in enumerate(range(0, n, step)):
            c1 = min(c0 + step, n)
            call("ps_hash_bwd_lm", ptr(x01[c0 * S:c1 * S]), (c1 - c0) * S, None, host_floats(sc), L, F, log2T, ptr(douts[i]), ptr(dtable), None, stream())
    res = []
    for fn in (fwd, bwd):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): fn()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 3)
    print(f"slices {k:2d}: gather {res[0]:.3f} ms   scatter {res[1]:.3f} ms", flush=True)
