#!/usr/bin/env python
"""Micro-benchmark of the hash-grid kernels: sweep features/level, table size and point coherence.
    python tools/bench_hash.py [--points 4194304]
Prints one line per configuration: fwd / bwd ms and algorithmic GB/s (SURVEY §8d byte counts)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from presight_b200 import ops, synthetic  # noqa: E402
from presight_b200.field_components.encodings import HashEncoding  # noqa: E402


def ray_points(P, S, dev):
    """P points as P/S rays x S consecutive samples (coherent along the ray), mapped into the unit cube."""
    n = P // S
    g = torch.Generator(device="cpu").manual_seed(1)
    o = torch.rand(n, 1, 3, generator=g) * 0.2 + 0.4
    d = torch.nn.functional.normalize(torch.randn(n, 1, 3, generator=g), dim=-1)
    t = torch.sort(torch.rand(n, S, 1, generator=g) ** 2, dim=1).values * 0.4
    return (o + d * t).clamp(0.001, 0.999).reshape(-1, 3).to(dev)


def time_it(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=65536 * 64)
    args = ap.parse_args()
    dev = "cuda"
    P = args.points
    pts = {"uniform": torch.rand(P, 3, device=dev), "rays64": ray_points(P, 64, dev)}
    print(f"{'config':34s} {'points':8s} {'fwd ms':>8s} {'fwd GB/s':>9s} {'bwd ms':>8s} {'bwd GB/s':>9s} {'Gelem/s':>8s}")
    for (L, F, log2T, hi) in [(16, 2, 22, 2048), (16, 2, 21, 2048), (16, 2, 20, 2048), (16, 2, 19, 2048), (16, 1, 22, 2048), (16, 1, 19, 2048),
                              (8, 1, 20, 4096), (8, 2, 20, 4096), (10, 4, 20, 16384), (16, 2, 16, 2048)]:
        enc = HashEncoding(num_levels=L, min_res=16, max_res=hi, log2_hashmap_size=log2T, features_per_level=F).to(dev)
        table = enc.hash_table.detach()
        dtable = torch.zeros_like(table)
        for name, x in pts.items():
            out = torch.empty(P, L * F, device=dev)
            dout = torch.randn(P, L * F, device=dev)
            sc = enc._scalings_host

            def fwd():
                ops.call("ps_hash_fwd", ops.ptr(x), P, ops.ptr(table), ops.host_floats(sc), L, F, log2T, ops.ptr(out),
                         ops.stream())

            def bwd():
                ops.call("ps_hash_bwd", ops.ptr(x), P, None, ops.host_floats(sc), L, F, log2T, ops.ptr(dout),
                         ops.ptr(dtable), None, ops.stream())
            def fwd_lm():
                ops.call("ps_hash_fwd_lm", ops.ptr(x), P, ops.ptr(table), ops.host_floats(sc), L, F, log2T, ops.ptr(out),
                         ops.stream())

            def bwd_lm():
                ops.call("ps_hash_bwd_lm", ops.ptr(x), P, None, ops.host_floats(sc), L, F, log2T, ops.ptr(dout),
                         ops.ptr(dtable), None, ops.stream())
            if os.environ.get("LM") == "1":
                fwd, bwd = fwd_lm, bwd_lm
            tf, tb = time_it(fwd), time_it(bwd)
            bf, bb = synthetic.hash_bytes_fwd(L, F) * P, synthetic.hash_bytes_bwd(L, F) * P
            print(f"L{L} F{F} T2^{log2T} {name:18s} {P:8d} {tf:8.3f} {bf/tf/1e6:9.0f} {tb:8.3f} {bb/tb/1e6:9.0f} "
                  f"{P*L*8*F/tb/1e6:8.1f}")
        del enc, table, dtable


if __name__ == "__main__":
    main()
