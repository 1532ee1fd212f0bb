#!/usr/bin/env python
"""Data-parallel formulation of the z-anti-aliased interlevel loss — the plan for a warp-per-ray kernel.

The shipped kernel (csrc/zaa_core.h) runs the reference's algorithm sequentially, one thread per ray: a two-way merge, two
nested running sums, a forward sweep over the sorted queries.  Every one of those steps has a formulation without
loop-carried control flow, which is what a warp needs:

  merge            rank of knot A[e] = c[e] - r  is  e + #{k : B[k] <  A[e]}       (B[k] = c[k] + r; "A first on ties")
                   rank of knot B[k]             is  k + #{e : A[e] <= B[k]}       -> two binary searches per knot
  running sums     inclusive scans (fp64 carry, fp32 outputs) over the K = 2S + 2 knots in rank order
  interval lookup  j  = #{knots <= x} - 1                                           -> binary search per query
                   i0 = first knot of the flat run of the integral ending at j     -> binary search (cdf, side = left)
                   right density = yr[0] if the integral is flat from j + 1 to the end else yr[j + 1]

This file states that formulation in numpy (vectorised over rays) and checks it against the live reference's fixture
(tests/golden/zaa.npz): `python tools/zaa_parallel_prototype.py`; tests/test_zaa_host.py runs the same check.
"""
import os
import sys

import numpy as np

F32 = np.float32


def zaa_level(c, w, cp, wp, r):
    """c [N,S+1], w [N,S], cp [N,Sp+1], wp [N,Sp] float32 -> (w_s [N,Sp], loss terms [N,Sp], d terms / d wp)."""
    N, S = w.shape
    K = 2 * S + 2
    rf, two_r = F32(r), F32(2.0 * r)
    A, B = (c - rf).astype(F32), (c + rf).astype(F32)
    wn = (w / (c[:, 1:] - c[:, :-1])).astype(F32)
    pad = np.zeros((N, 1), F32)
    y1 = ((np.concatenate([wn, pad], 1) - np.concatenate([pad, wn], 1)) / two_r).astype(F32)        # [N,S+1]
    xr, y2 = np.empty((N, K), F32), np.empty((N, K), F32)
    e = np.arange(S + 1)
    for n in range(N):                                      # (per ray: what one warp does)
        pos_a = e + np.searchsorted(B[n], A[n], side="left")            # #{B < A[e]}
        pos_b = e + np.searchsorted(A[n], B[n], side="right")           # #{A <= B[k]}
        xr[n, pos_a], xr[n, pos_b] = A[n], B[n]
        y2[n, pos_a], y2[n, pos_b] = y1[n], -y1[n]
    inner = np.cumsum(y2[:, :-1].astype(np.float64), 1).astype(F32)                                  # last knot dropped
    prod = ((xr[:, 1:] - xr[:, :-1]) * inner).astype(F32)
    yr = np.concatenate([pad, np.maximum(np.cumsum(prod.astype(np.float64), 1).astype(F32), 0)], 1)  # [N,K]
    area = ((F32(0.5) * (yr[:, 1:] + yr[:, :-1])).astype(F32) * (xr[:, 1:] - xr[:, :-1])).astype(F32)
    cdf = np.concatenate([pad, np.cumsum(area.astype(np.float64), 1).astype(F32)], 1)                # [N,K]
    ret = np.empty_like(cp)
    for n in range(N):
        x = cp[n]
        j = np.searchsorted(xr[n], x, side="right") - 1
        left = j < 0
        j = np.clip(j, 0, K - 1)
        i0 = np.searchsorted(cdf[n], cdf[n, j], side="left")             # first knot of the flat run
        jn = np.minimum(j + 1, K - 1)
        last = j == K - 1
        f0 = yr[n, i0]
        f1 = np.where(last | (cdf[n, K - 1] == cdf[n, jn]), yr[n, 0], yr[n, jn])
        x0 = xr[n, j]
        with np.errstate(divide="ignore", invalid="ignore"):
            o = ((x - x0) / (xr[n, jn] - x0)).astype(F32)
        o = np.where(last, np.where(x > x0, F32(1), F32(0)), np.clip(np.nan_to_num(o, nan=0.0), 0, 1)).astype(F32)
        val = (cdf[n, j] + ((x - x0) * ((f0 + f1 * o).astype(F32) + (f0 * (F32(1) - o)).astype(F32)).astype(F32)).astype(F32)
               / F32(2)).astype(F32)
        ret[n] = np.where(left, F32(0), val)
    w_s = (ret[:, 1:] - ret[:, :-1]).astype(F32)
    rr = np.maximum(w_s - wp, 0).astype(F32)
    den = (wp + F32(1e-5)).astype(F32)
    return w_s, rr * rr / den, -2 * rr / den - rr * rr / (den * den)


def check(fixture_path):
    z = np.load(fixture_path)
    pulse = [float(v) for v in z["pulse_width"]]
    worst = 0.0
    for case in "abc":
        total = 0.0
        for i in range(int(z[f"{case}/n_levels"])):
            c, w, cp, wp = (np.ascontiguousarray(z[f"{case}/{k}"], dtype=F32) for k in ("c", "w", f"t{i}", f"w{i}"))
            w_s, terms, grad = zaa_level(c, w, cp, wp, pulse[i])
            ref_ws, ref_g = z[f"{case}/ws{i}"], z[f"{case}/g{i}"]
            e_ws = np.abs(w_s - ref_ws).max() / np.abs(ref_ws).max()
            e_g = np.abs(grad / terms.size - ref_g).max() / np.abs(ref_g).max()
            worst = max(worst, e_ws, e_g)
            total += terms.astype(np.float64).mean()
        e_l = abs(total - float(z[f"{case}/loss"])) / abs(float(z[f"{case}/loss"]))
        worst = max(worst, e_l)
        print(f"case {case}: loss rel err {e_l:.2e}")
    return worst


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    worst = check(os.path.join(os.path.dirname(here), "tests", "golden", "zaa.npz"))
    print(f"worst scale-relative error over w_s / gradients / losses: {worst:.2e}")
    sys.exit(0 if worst < 2e-5 else 1)
