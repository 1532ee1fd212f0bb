"""Ray containers (reference: nerfstudio/cameras/rays.py).

Light-weight counterparts of `Frustums`, `RaySamples` and `RayBundle` with the same field names and
methods on the hot path (`get_positions`, `get_weights`, `get_ray_samples`).  Tensors keep the
reference's shapes: per-ray fields are [N,1,C] inside a RaySamples, per-sample fields are [N,S,1].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import torch
from torch import Tensor

from .. import ops


@dataclass
class Frustums:
    origins: Tensor        # [N,1,3] or [N,S,3]
    directions: Tensor     # [N,1,3] or [N,S,3]
    starts: Tensor         # [N,S,1]
    ends: Tensor           # [N,S,1]
    pixel_area: Optional[Tensor] = None
    offsets: Optional[Tensor] = None
    eu_bins: Optional[Tensor] = None   # [N,S+1] euclidean bin edges when the samples are contiguous bins

    @property
    def shape(self):
        return self.starts.shape[:-1]

    def get_positions(self) -> Tensor:
        """origins + directions * (starts + ends) / 2 (rays.py:49-58) via `ps_sample_positions`."""
        if self.eu_bins is not None and self.origins.shape[-2] == 1 and self.offsets is None:
            return ops.sample_positions(self.origins[:, 0], self.directions[:, 0], self.eu_bins)
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None
    deltas: Optional[Tensor] = None
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None
    sp_bins: Optional[Tensor] = None   # [N,S+1] spacing-domain bin edges (contiguous bins)

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights(self, densities: Tensor) -> Tensor:
        """alpha-compositing weights (rays.py:128-150) via `ps_weights_fwd` / `ps_weights_bwd`."""
        N, S = densities.shape[0], densities.shape[1]
        w = ops.get_weights(self.deltas.reshape(N, S), densities.reshape(N, S))
        return w.view(N, S, 1)


@dataclass
class RayBundle:
    origins: Tensor                 # [N,3]
    directions: Tensor              # [N,3]
    pixel_area: Optional[Tensor] = None
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None  # [N,1]
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self) -> int:
        return self.origins.shape[0]

    def get_ray_samples(self, bin_starts: Tensor, bin_ends: Tensor, spacing_starts: Optional[Tensor] = None,
                        spacing_ends: Optional[Tensor] = None, spacing_to_euclidean_fn: Optional[Callable] = None,
                        eu_bins: Optional[Tensor] = None, sp_bins: Optional[Tensor] = None) -> RaySamples:
        """rays.py:251-295."""
        deltas = bin_ends - bin_starts
        frustums = Frustums(origins=self.origins[:, None, :], directions=self.directions[:, None, :],
                            starts=bin_starts, ends=bin_ends,
                            pixel_area=None if self.pixel_area is None else self.pixel_area[:, None, :],
                            eu_bins=eu_bins)
        return RaySamples(frustums=frustums,
                          camera_indices=None if self.camera_indices is None else self.camera_indices[:, None, :],
                          deltas=deltas, spacing_starts=spacing_starts, spacing_ends=spacing_ends,
                          spacing_to_euclidean_fn=spacing_to_euclidean_fn,
                          metadata={k: v[:, None, :] for k, v in self.metadata.items()} if self.metadata else None,
                          times=None if self.times is None else self.times[:, None, :], sp_bins=sp_bins)

    @staticmethod
    def samples_from_bins(bundle: "RayBundle", sp_bins: Tensor, eu_bins: Tensor, fn: Optional[Callable]) -> RaySamples:
        return bundle.get_ray_samples(bin_starts=eu_bins[..., :-1, None], bin_ends=eu_bins[..., 1:, None],
                                      spacing_starts=sp_bins[..., :-1, None], spacing_ends=sp_bins[..., 1:, None],
                                      spacing_to_euclidean_fn=fn, eu_bins=eu_bins, sp_bins=sp_bins)
