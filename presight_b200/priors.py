"""Prior post-processing and wire format (reference: nerfstudio/scripts/extract_priors.py:156-208, 216-245; reader:
online-mapping/plugin/datasets/prior_utils/city_prior.py:59-73).

The reference concatenates every hit point of a tile on the HOST (~300 GB of RAM, docs/building_priors.md:65), filters
by density, voxel-down-samples with open3d and traces features per voxel in a Python loop.  Here the hit points stay in
HBM and go through a GPU hash-voxelizer (csrc/voxelize.cu): per-voxel centre of mass, colour mean, fp64 feature mean
rounded to fp16, hit count, then the hit-quantile filter — and the result is written as the same pickled dict
`{points f32, features f16, colors f32, hits, origin f32}` the occupancy / online-mapping plugins read.

`PriorVoxelizer` is the streaming form (one `add()` per camera / chunk once the bound is known); `postprocess_priors`
is the one-shot form over resident arrays.  CUDA tensors only — no CPU fallback (the CPU restatement is oracle/).
"""
from __future__ import annotations

import pickle
from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

from ._lib import call, ptr, stream


def _next_pow2(n: int) -> int:
    return 1 << max(1, (int(n) - 1).bit_length())


class PriorVoxelizer:
    """Accumulates hit points into voxels of `voxel_size` metres.  The voxel lattice is anchored, as open3d anchors it, at
    (min over all density-filtered points) - 1 - voxel_size / 2, so the bound must be known before the first `add`:
    pass `min_point` (e.g. from a first pass with `min_bound`) or use `postprocess_priors` on resident arrays."""

    def __init__(self, min_point: Tensor, voxel_size: float = 0.4, feature_dim: int = 64, capacity: int = 1 << 22,
                 with_colors: bool = True) -> None:
        assert min_point.is_cuda and min_point.numel() == 3, "min_point: 3 floats on the GPU"
        dev = min_point.device
        self.dev, self.voxel_size, self.C = dev, float(voxel_size), int(feature_dim)
        self.min_point = min_point.detach().to(torch.float32).contiguous()
        self.capacity = _next_pow2(capacity)
        self.keys = torch.full((self.capacity,), -1, device=dev, dtype=torch.int64)
        self.counts = torch.zeros(self.capacity, device=dev, dtype=torch.int32)        # u32 on the device
        self.sum_xyz = torch.zeros(self.capacity, 3, device=dev, dtype=torch.float64)
        self.sum_col = torch.zeros(self.capacity, 3, device=dev, dtype=torch.float64) if with_colors else None
        self.sum_feat = torch.zeros(self.capacity, self.C, device=dev, dtype=torch.float64) if self.C else None
        self.status = torch.zeros(1, device=dev, dtype=torch.int32)

    @staticmethod
    def min_bound(points: Tensor, densities: Optional[Tensor] = None, running: Optional[Tensor] = None) -> Tensor:
        """Per-axis min over the points with density > 1 (extract_priors.py:157, 236), folded into `running` if given."""
        out = running if running is not None else torch.full((3,), float("inf"), device=points.device)
        p = points.detach().to(torch.float32).contiguous()
        d = None if densities is None else densities.detach().to(torch.float32).contiguous()
        call("ps_voxel_min_bound", ptr(p), ptr(d), p.shape[0], ptr(out), stream())
        return out

    def add(self, points: Tensor, features: Optional[Tensor], colors: Optional[Tensor],
            densities: Optional[Tensor] = None) -> None:
        """points [n,3] f32 metres, features [n,C] f16, colors [n,3] f32, densities [n] (None = keep all)."""
        p = points.detach().to(torch.float32).contiguous()
        f = None if self.sum_feat is None else features.detach().to(torch.float16).contiguous()
        c = None if self.sum_col is None else colors.detach().to(torch.float32).contiguous()
        d = None if densities is None else densities.detach().to(torch.float32).contiguous()
        assert f is None or f.shape == (p.shape[0], self.C)
        call("ps_voxel_accumulate", ptr(p), ptr(f), ptr(c), ptr(d), p.shape[0], self.C, ptr(self.min_point),
             self.voxel_size, ptr(self.keys), self.capacity, ptr(self.counts), ptr(self.sum_xyz), ptr(self.sum_col),
             ptr(self.sum_feat), ptr(self.status), stream())

    def finalize(self, hit_thr_ratio: float = 0.2) -> Dict[str, Tensor]:
        """-> {"points" [M,3] f32, "features" [M,C] f16, "colors" [M,3] f32, "hits" [M] i64, "hit_thr" f64 scalar,
        "n_voxels" int} on the device, voxels in ascending (ix, iy, iz) order, after the `hits > quantile` filter
        (extract_priors.py:188-196)."""
        st = int(self.status.item())
        if st & 1:
            raise RuntimeError(f"PriorVoxelizer: hash table of {self.capacity} slots is full — construct it with a larger capacity")
        if st & 2:
            raise RuntimeError("PriorVoxelizer: a voxel index exceeds 21 bits per axis (tile larger than 2^21 voxels across)")
        dev = self.dev
        n_out = torch.zeros(1, device=dev, dtype=torch.int64)
        cap = self.capacity
        okeys = torch.empty(cap, device=dev, dtype=torch.int64)
        oxyz = torch.empty(cap, 3, device=dev, dtype=torch.float32)
        ocol = torch.empty(cap, 3, device=dev, dtype=torch.float32) if self.sum_col is not None else None
        ofeat = torch.empty(cap, self.C, device=dev, dtype=torch.float16) if self.sum_feat is not None else None
        ohits = torch.empty(cap, device=dev, dtype=torch.int64)
        call("ps_voxel_finalize", ptr(self.keys), cap, ptr(self.counts), ptr(self.sum_xyz), ptr(self.sum_col),
             ptr(self.sum_feat), self.C, ptr(n_out), ptr(okeys), ptr(oxyz), ptr(ocol), ptr(ofeat), ptr(ohits), stream())
        m = int(n_out.item())
        if m == 0:
            raise RuntimeError("PriorVoxelizer: no points passed the density filter")
        order = torch.argsort(okeys[:m])                    # canonical voxel order (open3d's is a hash map's; see oracle/)
        hits = ohits[:m][order]
        bins = int(hits.max().item()) + 1
        hist = torch.zeros(bins, device=dev, dtype=torch.int32)
        thr = torch.zeros(1, device=dev, dtype=torch.float64)
        status = torch.zeros(1, device=dev, dtype=torch.int32)
        call("ps_hits_quantile", ptr(hits), m, float(hit_thr_ratio), ptr(hist), bins, ptr(thr), ptr(status), stream())
        keep = hits.to(torch.float64) > thr                 # extract_priors.py:191
        sel = order[keep]
        return {"points": oxyz[:m][sel], "features": None if ofeat is None else ofeat[:m][sel],
                "colors": None if ocol is None else ocol[:m][sel], "hits": hits[keep], "hit_thr": thr[0], "n_voxels": m}


def postprocess_priors(points: Tensor, features: Tensor, colors: Tensor, densities: Optional[Tensor],
                       voxel_size: float = 0.4, hit_thr_ratio: float = 0.2, capacity: Optional[int] = None) -> Dict[str, Tensor]:
    """extract_priors.py:156-196 on resident arrays (one tile): density filter, voxel down-sampling, per-voxel means,
    hit-quantile filter.  `capacity`: hash-table slots (default: 2 x the number of points, rounded up to a power of 2)."""
    mn = PriorVoxelizer.min_bound(points, densities)
    vox = PriorVoxelizer(mn, voxel_size, features.shape[1], capacity or 2 * max(points.shape[0], 1024))
    vox.add(points, features, colors, densities)
    return vox.finalize(hit_thr_ratio)


def priors_to_dict(result: Dict[str, Tensor], origin) -> Dict[str, np.ndarray]:
    """The dict extract_priors.py:199-208 pickles (host arrays, reference dtypes)."""
    return {"points": result["points"].cpu().numpy().astype(np.float32),
            "features": result["features"].cpu().numpy().astype(np.float16),
            "colors": result["colors"].cpu().numpy().astype(np.float32),
            "hits": result["hits"].cpu().numpy(),
            "origin": np.asarray(origin.cpu() if torch.is_tensor(origin) else origin).astype(np.float32)}


def save_priors(path: str, result: Dict[str, Tensor], origin) -> None:
    """Write `extracted_priors.pkl` (extract_priors.py:197-208)."""
    with open(path, "wb") as f:
        pickle.dump(priors_to_dict(result, origin), f)
