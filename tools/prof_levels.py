#!/usr/bin/env python
"""Run the fused proposal / field level kernels once or a few times at C2 sizes (for ncu captures and quick timings).

    python tools/prof_levels.py [--rays 65536] [--iters 3] [--what prop,field]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from presight_b200 import fused, ops
import numpy as np


def scalings(L, lo, hi):
    g = np.exp((np.log(hi) - np.log(lo)) / (L - 1)) if L > 1 else 1.0
    return tuple(float(np.floor(lo * g ** l)) for l in range(L))


def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    if iters <= 0:
        return float("nan")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--what", default="prop,field")
    args = ap.parse_args()
    n, dev = args.rays, "cuda"
    g = torch.Generator().manual_seed(0)
    from presight_b200 import synthetic
    rays = synthetic.make_rays(n, seed=1)
    o, d = rays["origins"].to(dev), rays["directions"].to(dev)
    aabb = [float(v) for v in synthetic.tile_aabb().reshape(-1)]
    ops.PROBE = ops.KernelProbe()
    if "prop" in args.what:
        for S in (128, 64):
            L, F, log2T, H = 8, 1, 20, 64
            grid = fused.GridMeta(scalings(L, 16, 1024 if S == 128 else 4096), log2T, F)
            table = ((torch.rand(L << log2T, F, generator=g) * 2 - 1) * 1e-3).to(dev).requires_grad_(True)
            ws = [(torch.randn(H, L * F, generator=g) / (L * F) ** 0.5).to(dev).requires_grad_(True),
                  (torch.randn(1, H, generator=g) / H ** 0.5).to(dev).requires_grad_(True)]
            bs = [torch.zeros(H, device=dev, requires_grad=True), torch.zeros(1, device=dev, requires_grad=True)]
            eu = (torch.rand(n, S + 1, generator=g) * (40.0 / S) + 0.001).cumsum(-1).to(dev)
            gw = torch.randn(n, S, 1, generator=g).to(dev) * 1e-3

            def run():
                w = fused._PropLevelTc5.apply(o, d, eu, table, aabb, True, grid, None, ws[0], bs[0], ws[1], bs[1])
                torch.autograd.grad((w * gw).sum(), [table, *ws, *bs])
            print(f"prop S={S}: {timeit(run, args.iters):.3f} ms fwd+bwd", flush=True)
    if "field" in args.what:
        S, L, F, log2T, A = 64, 16, 2, 22, 16
        grid = fused.GridMeta(scalings(L, 16, 2048), log2T, F)
        table = ((torch.rand(L << log2T, F, generator=g) * 2 - 1) * 1e-3).to(dev).requires_grad_(True)
        dims = [(L * F, 64, 80), (64, 64, 64, 64), (31 + A, 64, 64, 3)]
        ws, bs = [], []
        for dd in dims:
            for i in range(len(dd) - 1):
                ws.append((torch.randn(dd[i + 1], dd[i], generator=g) / dd[i] ** 0.5).to(dev).requires_grad_(True))
                bs.append(torch.zeros(dd[i + 1], device=dev, requires_grad=True))
        eu = (torch.rand(n, S + 1, generator=g) * (40.0 / S) + 0.001).cumsum(-1).to(dev)
        app = torch.randn(n, A, generator=g).to(dev).requires_grad_(True)

        def run():
            out = fused._FieldLevelTc5.apply(o, d, eu, app, table, aabb, True, grid, 0.5, *ws, *bs)
            loss = out[1].sum() + out[5].sum() * 0.1 + out[0].sum() * 0.01
            torch.autograd.grad(loss, [table, app, *ws, *bs])
        print(f"field S={S}: {timeit(run, args.iters):.3f} ms fwd+bwd", flush=True)
    for k, v in sorted(ops.PROBE.summary().items()):
        print(f"  {k:28s} {v[1]:8.3f} ms x{v[0]}")


if __name__ == "__main__":
    main()
