"""PropNetDensityField (reference: nerfstudio/fields/PreSight/prop_density_field.py) on the b200 kernels."""
from __future__ import annotations

from copy import deepcopy
from typing import Literal, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..cameras.rays import RaySamples
from ..field_components.encodings import HashEncoding
from ..field_components.mlp import MLP
from ..field_components.spatial_distortions import SpatialDistortion
from .base_field import Field


class PropNetDensityField(Field):
    def __init__(
        self,
        aabb: Tensor,
        num_layers: int = 2,
        hidden_dim: int = 64,
        spatial_distortion: Optional[SpatialDistortion] = None,
        use_linear: bool = False,
        num_levels: int = 8,
        max_res: int = 1024,
        base_res: int = 16,
        log2_hashmap_size: int = 18,
        features_per_level: int = 2,
        implementation: Literal["b200", "b200+fp32"] = "b200",
        field_type: Literal["iNGP", "TriPlane"] = "iNGP",
    ) -> None:
        super().__init__()
        self.register_buffer("aabb", deepcopy(aabb))
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        if field_type != "iNGP":
            raise ValueError(f"Unknown `field_type`: {field_type}")
        self.encoding = HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res,
                                     log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level,
                                     implementation=implementation)
        self._precision = ops.PREC_BF16 if implementation == "b200" else ops.PREC_TF32X3
        if not self.use_linear:
            network = MLP(in_dim=self.encoding.get_out_dim(), num_layers=num_layers, layer_width=hidden_dim, out_dim=1,
                          activation=nn.ReLU(), out_activation=None, implementation=implementation)
            self.mlp_base = torch.nn.Sequential(self.encoding, network)
        else:
            self.linear = torch.nn.Linear(self.encoding.get_out_dim(), 1)

    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, None]:
        return self.density_fn(ray_samples.frustums.get_positions()), None

    def density_fn(self, positions: Tensor) -> Tensor:
        """prop_density_field.py:129-153."""
        x01, selector = ops.normalize_positions(positions, self.aabb_host(), self.spatial_distortion is not None)
        flat = x01.view(-1, 3)
        if not self.use_linear:
            raw = self.mlp_base(flat)
        else:
            raw = ops.mlp(self.encoding(flat), [self.linear.weight], [self.linear.bias], ops.ACT_NONE, self._precision)
        return ops.trunc_exp(raw, selector.reshape(-1)).view(*positions.shape[:-1], 1)

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None) -> dict:
        return {}

    def level_weights(self, origins: Tensor, directions: Tensor, eu_bins: Tensor) -> Tensor:
        """Fast path of `get_weights(density_fn(positions))` for contiguous bins: one autograd node, a fixed chain
        of kernels (ray_points -> hash -> MLP with density epilogue -> weights), no intermediate torch ops."""
        from .. import fused
        enc = self.encoding
        if self.use_linear:
            layers = [self.linear]
        else:
            layers = list(self.mlp_base[1].layers)
        dims = (layers[0].weight.shape[1],) + tuple(l.weight.shape[0] for l in layers)
        return fused.prop_level_weights(
            origins, directions, eu_bins, enc.hash_table, self.aabb_host(), self.spatial_distortion is not None,
            fused.GridMeta(enc._scalings_host, enc.log2_hashmap_size, enc.features_per_level),
            fused.MlpMeta(dims, ops.ACT_NONE), self._precision, [l.weight for l in layers], [l.bias for l in layers])
