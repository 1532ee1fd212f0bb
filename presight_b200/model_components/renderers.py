"""Renderers (reference: nerfstudio/model_components/renderers.py) on the b200 compositing kernels."""
from __future__ import annotations

from typing import Literal, Optional, Union

import torch
from torch import Tensor, nn

from .. import ops
from ..cameras.rays import RaySamples

BackgroundColor = Union[Literal["random", "last_sample", "black", "white"], Tensor]


def _no_packed(ray_indices, num_rays):
    if ray_indices is not None or num_rays is not None:
        raise NotImplementedError("packed samples (nerfacc) are not on PreSight's path")


class RGBRenderer(nn.Module):
    """renderers.py:58-229."""

    def __init__(self, background_color: BackgroundColor = "random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb: Tensor, weights: Tensor, background_color: BackgroundColor = "random",
                    ray_indices=None, num_rays=None) -> Tensor:
        _no_packed(ray_indices, num_rays)
        N, S = weights.shape[0], weights.shape[1]
        w = weights.reshape(N, S)
        comp_rgb = ops.render(w, rgb)
        if isinstance(background_color, str) and background_color == "random":
            return comp_rgb
        acc = ops.render(w, None)
        if isinstance(background_color, str):
            if background_color == "last_sample":
                bg = rgb[..., -1, :]
            elif background_color == "black":
                bg = torch.zeros(3, device=rgb.device)
            elif background_color == "white":
                bg = torch.ones(3, device=rgb.device)
            else:
                raise ValueError(background_color)
        else:
            bg = background_color
        return comp_rgb + bg * (1.0 - acc)

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices=None, num_rays=None,
                background_color: Optional[BackgroundColor] = None) -> Tensor:
        if background_color is None:
            background_color = self.background_color
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        rgb = self.combine_rgb(rgb, weights, background_color=background_color, ray_indices=ray_indices,
                               num_rays=num_rays)
        if not self.training:
            rgb = torch.clamp(rgb, min=0.0, max=1.0)
        return rgb


class AccumulationRenderer(nn.Module):
    """renderers.py:286-314."""

    @classmethod
    def forward(cls, weights: Tensor, ray_indices=None, num_rays=None) -> Tensor:
        _no_packed(ray_indices, num_rays)
        return ops.render(weights.reshape(weights.shape[0], weights.shape[1]), None)


class DepthRenderer(nn.Module):
    """renderers.py:317-383 ("threshold" is PreSight's rename of "median")."""

    def __init__(self, method: Literal["threshold", "expected"] = "threshold") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights: Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None,
                threshold: float = 0.5) -> Tensor:
        _no_packed(ray_indices, num_rays)
        N, S = weights.shape[0], weights.shape[1]
        w = weights.reshape(N, S)
        if self.method == "threshold":
            eu = ray_samples.frustums.eu_bins
            if eu is None:
                eu = torch.cat([ray_samples.frustums.starts[..., 0], ray_samples.frustums.ends[..., -1:, 0]], dim=-1)
            depth, _ = ops.depth_threshold(w, eu, threshold)
            return depth
        if self.method == "expected":
            eps = 1e-10
            steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
            depth = ops.render(w, steps) / (ops.render(w, None) + eps)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(f"Method {self.method} not implemented")
