"""SkyField (reference: nerfstudio/fields/PreSight/sky_field.py) on the b200 kernels."""
from __future__ import annotations

from typing import Dict, Literal, Optional

import torch
from torch import Tensor, nn

from ..cameras.rays import RaySamples
from ..field_components.encodings import SHEncoding
from ..field_components.mlp import MLP
from .ingp_field import FieldHeadNames


class SkyField(nn.Module):
    def __init__(self, direction_encoding: str = "SHEncoding", mlp_num_layers: int = 3, mlp_layer_width: int = 64,
                 appearance_embedding_dim: int = 32, use_semantics: bool = False, semantic_dim: int = 64,
                 implementation: Literal["b200", "b200+fp32"] = "b200") -> None:
        super().__init__()
        self.use_semantics = use_semantics
        self.appearance_embedding_dim = appearance_embedding_dim
        if direction_encoding != "SHEncoding":
            raise ValueError(direction_encoding)
        self.direction_encoding = SHEncoding(levels=4, implementation=implementation)
        self.rgb_head = MLP(in_dim=self.direction_encoding.get_out_dim() + self.appearance_embedding_dim,
                            num_layers=mlp_num_layers, layer_width=mlp_layer_width, out_dim=3, activation=nn.ReLU(),
                            out_activation=nn.Sigmoid(), implementation=implementation)
        if self.use_semantics:
            self.semantic_head = MLP(in_dim=self.direction_encoding.get_out_dim(), num_layers=mlp_num_layers,
                                     layer_width=mlp_layer_width, out_dim=semantic_dim, activation=nn.ReLU(),
                                     out_activation=None, implementation=implementation)

    def get_outputs(self, directions: Tensor, appearance_embedding: Optional[Tensor]) -> Dict[str, Tensor]:
        """sky_field.py:95-111."""
        d = self.direction_encoding.forward_raw(directions)
        outputs = {}
        if appearance_embedding is not None:
            outputs[FieldHeadNames.RGB] = self.rgb_head(torch.cat([d, appearance_embedding], dim=-1))
        else:
            outputs[FieldHeadNames.RGB] = self.rgb_head(d)
        if self.use_semantics:
            outputs[FieldHeadNames.SEMANTICS] = self.semantic_head(d)
        return outputs

    def forward(self, ray_samples: RaySamples, appearance_embedding: Optional[Tensor]) -> Dict[str, Tensor]:
        directions = ray_samples.frustums.directions[:, 0, :].contiguous()
        if appearance_embedding is not None and appearance_embedding.dim() == 3:   # [N,S,A] (reference) or [N,A]
            appearance_embedding = appearance_embedding[:, 0, :]
        if appearance_embedding is not None:
            appearance_embedding = appearance_embedding.contiguous()
        return self.get_outputs(directions, appearance_embedding)
