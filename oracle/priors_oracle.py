"""CPU restatement (numpy) of the prior post-processing of scripts/extract_priors.py — TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, for the accumulated hit points of one tile (XP = nerfstudio/scripts/extract_priors.py):
  XP:156-165   density filter  `all_hit_points_densities > 1.0`
  XP:167-173   `points_downsample_to_voxels` -> XP:216-245: open3d `PointCloud.voxel_down_sample_and_trace(voxel_size,
               min_bound = points.min(0) - 1, max_bound = points.max(0) + 1)`
  XP:175-191   per voxel: colour = float32 mean, feature = float64 mean of the fp16 features cast back to fp16,
               hits = number of points; `hit_thr = np.quantile(hits, hit_thr_ratio)`; keep `hits > hit_thr`
  XP:199-208   the pickled dict {points f32, features f16, colors f32, hits, origin f32}
and the reader the perception plugins use (online-mapping/plugin/datasets/prior_utils/city_prior.py:59-73).

Third-party arithmetic: open3d is NOT vendored in /root/reference and NOT pinned (occupancy/requirements/optional.txt
lists a bare `open3d`; docs mention 0.9.0), and it is absent from this image, so the grouping step is restated from
open3d's published algorithm (cpp/open3d/geometry/PointCloud.cpp, `PointCloud::VoxelDownSampleAndTrace`, unchanged
between 0.9 and 0.18):
    voxel_min_bound = min_bound - voxel_size * 0.5                       (doubles; points are converted to double)
    voxel_index     = floor((point - voxel_min_bound) / voxel_size)      (Eigen::Vector3i)
    per voxel: running double sum of the points, count, list of point indices;
    output point = sum / count
open3d emits the voxels in the iteration order of a std::unordered_map (unspecified); every consumer of the pickle is
order-independent (VoxelizePriorPoints shuffles), so this restatement — and the product — emit voxels in ascending
(ix, iy, iz) order.  Parity status of the grouping: UNPINNED (no open3d here to generate a fixture from); everything
after it (means, quantile, selection, dict) is pinned by running the reference's own lines XP:175-208 on top of this
grouping in tests/golden/make_golden_priors.py.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

__all__ = ["voxel_down_sample_and_trace", "postprocess_priors", "read_priors_like_city_prior", "voxel_keys"]


def voxel_keys(points: np.ndarray, voxel_size: float) -> Tuple[np.ndarray, np.ndarray]:
    """-> (integer voxel indices [N,3] int64, voxel_min_bound [3] float64) for float32 points [N,3], with the bounds of
    XP:236-237 (computed in float32 like numpy does there) and open3d's half-voxel shift."""
    points = np.asarray(points)
    min_bound = (points.min(axis=0).reshape(3, 1) - 1.0).astype(np.float64).reshape(3)       # float32 arithmetic, then double
    vmb = min_bound - voxel_size * 0.5
    ref = (points.astype(np.float64) - vmb[None, :]) / voxel_size
    return np.floor(ref).astype(np.int64), vmb


def voxel_down_sample_and_trace(points: np.ndarray, voxel_size: float):
    """-> (voxel centres of mass [M,3] float64, list of index arrays, voxel indices [M,3]) in ascending voxel order."""
    idx, _ = voxel_keys(points, voxel_size)
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))
    sidx = idx[order]
    new = np.ones(len(order), dtype=bool)
    new[1:] = np.any(sidx[1:] != sidx[:-1], axis=1)
    starts = np.flatnonzero(new)
    ends = np.append(starts[1:], len(order))
    p64 = points.astype(np.float64)
    centres = np.stack([p64[order[s:e]].sum(axis=0) / (e - s) for s, e in zip(starts, ends)]) if len(starts) else \
        np.zeros((0, 3))
    traces = [np.sort(order[s:e]) for s, e in zip(starts, ends)]
    return centres, traces, sidx[starts]


def postprocess_priors(points: np.ndarray, features: np.ndarray, colors: np.ndarray, densities: Optional[np.ndarray],
                       origin: np.ndarray, voxel_size: float = 0.4, hit_thr_ratio: float = 0.2) -> Dict[str, np.ndarray]:
    """XP:156-208 on the concatenated hit points of a tile.  points [N,3] f32 (metres), features [N,C] f16,
    colors [N,3] f32, densities [N] f32 (None = already filtered)."""
    if densities is not None:
        sel = densities > 1.0                                               # XP:157
        points, features, colors = points[sel], features[sel], colors[sel]
    ds_points, ds_indices, _ = voxel_down_sample_and_trace(points, voxel_size)
    cols, feats, hits = [], [], []
    for indices in ds_indices:                                              # XP:178-186
        cols.append(colors[indices].mean(axis=0))
        feats.append(features[indices].astype(np.float64).mean(axis=0).astype(np.float16))
        hits.append(len(indices))
    hits = np.asarray(hits)
    cols, feats = np.stack(cols), np.stack(feats)
    hit_thr = np.quantile(np.asarray(hits), hit_thr_ratio)                   # XP:190
    keep = hits > hit_thr
    return {"points": ds_points[keep].astype(np.float32), "features": feats[keep].astype(np.float16),
            "colors": cols[keep].astype(np.float32), "hits": hits[keep], "origin": np.asarray(origin).astype(np.float32)}


def read_priors_like_city_prior(p: Dict[str, np.ndarray]):
    """What online-mapping/plugin/datasets/prior_utils/city_prior.py:63-73 does with one pickle."""
    xyz = p["points"].astype(np.float32) + p["origin"].astype(np.float32)
    xyz[:, 0:2] = -xyz[:, 0:2]
    hits = p["hits"].astype(np.float32)
    hits = hits / hits.mean()
    return xyz, p["features"].astype(np.float16), hits[:, None]
