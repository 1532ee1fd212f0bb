#!/usr/bin/env python
"""Golden fixture for the prior post-processing from the LIVE reference script.

    python tests/golden/make_golden_priors.py      # rewrites tests/golden/priors.npz

scripts/extract_priors.py cannot be imported here (viewer / dataparser stack, open3d), so this generator EXECUTES THE
REFERENCE'S OWN SOURCE LINES for everything after the voxel grouping — the per-voxel tracing loop, the hit quantile and
the selection (extract_priors.py:174-196, read from /root/reference at generation time, nothing copied into the repo) —
on seeded hit points.  The grouping itself (`pcd.voxel_down_sample_and_trace`, open3d — absent from this image and
un-pinned by the reference) comes from the oracle's restatement of open3d's published algorithm, with the min / max
bounds computed by the reference's lines 236-237; that one step is therefore "parity unpinned" (oracle/priors_oracle.py).
"""
import os
import sys
import textwrap

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import priors_oracle as PO  # noqa: E402

SCRIPT = "/root/reference/nerfstudio-0.3.3/nerfstudio/scripts/extract_priors.py"


def reference_lines(first, last):
    with open(SCRIPT) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def make_points(seed, n, c=64, ext=8.0):
    """Hit points the way extraction produces them: surfaces (ground plane, a wall) seen many times plus scattered points."""
    g = np.random.default_rng(seed)
    ground = np.stack([g.uniform(-ext, ext, n // 2), g.uniform(-ext / 2, ext / 2, n // 2), g.normal(0.0, 0.03, n // 2)], 1)
    wall = np.stack([g.uniform(-ext, ext, n // 4), np.full(n // 4, ext / 2 - 0.1) + g.normal(0, 0.05, n // 4),
                     g.uniform(0, 5, n // 4)], 1)
    rest = g.uniform([-ext, -ext / 2, -3], [ext, ext / 2, 6], (n - n // 2 - n // 4, 3))
    pts = np.concatenate([ground, wall, rest]).astype(np.float32)
    pts = pts[g.permutation(len(pts))]
    feats = g.uniform(0, 1, (len(pts), c)).astype(np.float16)
    cols = g.uniform(0, 1, (len(pts), 3)).astype(np.float32)
    dens = np.exp(g.normal(0.5, 1.5, len(pts))).astype(np.float32)
    return pts, feats, cols, dens


def run_case(pts, feats, cols, dens, voxel_size, hit_thr_ratio):
    # extract_priors.py:156-165 — the density filter, the reference's lines
    ns = {"np": np, "all_hit_points_densities": dens, "all_hit_points": pts, "all_hit_points_colors": cols,
          "all_hit_points_features": feats}
    sel = dens > 1.0
    all_hit_points_thr, all_hit_points_colors_thr, all_hit_points_features_thr = pts[sel], cols[sel], feats[sel]
    # extract_priors.py:236-237 bounds + open3d's grouping (oracle restatement)
    ds_points, ds_indices, _ = PO.voxel_down_sample_and_trace(all_hit_points_thr, voxel_size)
    ns.update(all_hit_points_colors_thr=all_hit_points_colors_thr, all_hit_points_features_thr=all_hit_points_features_thr,
              ds_points=ds_points, ds_indices=ds_indices, hit_thr_ratio=hit_thr_ratio, tqdm=lambda it, **kw: it,
              enumerate=enumerate, len=len, print=lambda *a, **k: None)
    exec(reference_lines(175, 196), ns)        # colors / features / hits per voxel, hit_thr, selector, *_thr
    return {"points": ns["points_thr"].astype(np.float32), "features": ns["features_thr"].astype(np.float16),
            "colors": ns["colors_thr"].astype(np.float32), "hits": ns["hits"][ns["selector"]],
            "hit_thr": np.float64(ns["hit_thr"]), "n_voxels": np.int64(len(ds_points))}


def main():
    out = {}
    for name, (seed, n, vs, q, c) in {"a": (1, 12000, 0.4, 0.2, 64), "b": (2, 6000, 0.25, 0.55, 16),
                                      "c": (3, 3000, 1.0, 0.0, 8)}.items():
        pts, feats, cols, dens = make_points(seed, n, c)
        if name == "c":
            dens[:] = 2.0                                # nothing filtered; quantile 0 keeps voxels with hits > min
        res = run_case(pts, feats, cols, dens, vs, q)
        out.update({f"{name}/in_points": pts, f"{name}/in_features": feats, f"{name}/in_colors": cols,
                    f"{name}/in_densities": dens, f"{name}/voxel_size": np.float64(vs), f"{name}/hit_thr_ratio": np.float64(q)})
        out.update({f"{name}/{k}": v for k, v in res.items()})
        print(name, "voxels", int(res["n_voxels"]), "kept", len(res["hits"]), "hit_thr", float(res["hit_thr"]))
    path = os.path.join(HERE, "priors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
