#!/usr/bin/env python
"""Per-level cost of the main-grid gather / scatter on C2-shaped sample points (one launch per level, L = 1).

    python tools/hash_levels.py [--rays 65536] [--samples 64]
Run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct`
to see where the DRAM traffic of the scatter comes from."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from presight_b200 import ops, synthetic, fused
from presight_b200._lib import call, ptr, stream, host_floats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--log2T", type=int, default=22)
    ap.add_argument("--F", type=int, default=2)
    ap.add_argument("--sort-res", type=int, default=0,
                    help="process the points in Morton order of a grid of this resolution (power of two) instead of ray "
                         "order: what a cell-sorted scatter of the coarse levels would see")
    args = ap.parse_args()
    dev, n, S, F, log2T = "cuda", args.rays, args.samples, args.F, args.log2T
    rays = synthetic.make_rays(n, seed=1)
    o, d = rays["origins"].to(dev), rays["directions"].to(dev)
    near, far, thr = 0.1 * 0.05, 1000 * 0.05, 100 * 0.05
    nears, fars = torch.full((n, 1), near, device=dev), torch.full((n, 1), far, device=dev)
    sp, eu = ops.spaced_bins(nears, fars, S, thr, torch.rand(n, 1, device=dev))
    aabb = [float(v) for v in synthetic.tile_aabb().reshape(-1)]
    x01, sel = fused._ray_points(o, d, eu.contiguous(), aabb, True)
    P = x01.shape[0]
    if args.sort_res:
        R = args.sort_res
        q = (x01.clamp(0, 1) * R).long().clamp_(0, R - 1)
        key = torch.zeros(P, dtype=torch.long, device=dev)
        for b in range(R.bit_length() - 1):
            for a in range(3):
                key |= ((q[:, a] >> b) & 1) << (3 * b + a)
        x01 = x01[torch.argsort(key)].contiguous()
        print(f"points in Morton order of a {R}^3 grid")
    L = 16
    g = np.exp((np.log(2048) - np.log(16)) / (L - 1))
    scal = [float(np.floor(16 * g ** l)) for l in range(L)]
    table = (torch.rand(1 << log2T, F, device=dev) * 2 - 1) * 1e-3
    dtable = torch.zeros_like(table)
    out = torch.empty(P * F, device=dev)
    dout = torch.randn(P * F, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def t(fn):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    print(f"{P} points, T=2^{log2T}, F={F}")
    tot_f = tot_b = 0.0
    for l, s in enumerate(scal):
        tf = t(lambda: call("ps_hash_fwd_lm", ptr(x01), P, ptr(table), host_floats([s]), 1, F, log2T, ptr(out), stream()))
        tb = t(lambda: call("ps_hash_bwd_lm", ptr(x01), P, None, host_floats([s]), 1, F, log2T, ptr(dout), ptr(dtable), None, stream()))
        tot_f += tf; tot_b += tb
        print(f"level {l:2d} res {int(s):5d}: fwd {tf:6.3f} ms   bwd {tb:6.3f} ms", flush=True)
    print(f"sum: fwd {tot_f:.3f} bwd {tot_b:.3f}")
    # all levels in one launch (what the model runs)
    table = (torch.rand(L << log2T, F, device=dev) * 2 - 1) * 1e-3
    dtable = torch.zeros_like(table)
    out = torch.empty(P * L * F, device=dev)
    dout = torch.randn(P * L * F, device=dev)
    for _ in range(3):
        tf = t(lambda: call("ps_hash_fwd_lm", ptr(x01), P, ptr(table), host_floats(scal), L, F, log2T, ptr(out), stream()))
        tb = t(lambda: call("ps_hash_bwd_lm", ptr(x01), P, None, host_floats(scal), L, F, log2T, ptr(dout), ptr(dtable), None, stream()))
    print(f"one launch, {L} levels: fwd {tf:.3f} ms  bwd {tb:.3f} ms")


if __name__ == "__main__":
    main()
