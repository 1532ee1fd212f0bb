"""Drop-in encodings (reference: nerfstudio/field_components/encodings.py).

`HashEncoding` keeps the reference constructor, attributes (`hash_table`, `scalings`, `hash_offset`,
`hash_table_size`, `tcnn_encoding`) and state-dict key (`hash_table`), but its forward/backward run the
sm_100a kernels (`ps_hash_fwd` / `ps_hash_bwd`) instead of `pytorch_fwd` (encodings.py:343-384).
"""
from __future__ import annotations

from typing import Literal, Optional

import numpy as np
import torch
from torch import Tensor, nn

from .. import ops

IMPLEMENTATIONS = ("b200", "b200+fp32")


class Encoding(nn.Module):
    def __init__(self, in_dim: int) -> None:
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")  # encodings.py:49-50
        super().__init__()
        self.in_dim = in_dim

    def get_out_dim(self) -> int:
        raise NotImplementedError


class HashEncoding(Encoding):
    """Multiresolution hash encoding (encodings.py:265-389) on hand-written CUDA kernels."""

    def __init__(
        self,
        num_levels: int = 16,
        min_res: int = 16,
        max_res: int = 1024,
        log2_hashmap_size: int = 19,
        features_per_level: int = 2,
        hash_init_scale: float = 0.001,
        implementation: Literal["b200", "b200+fp32"] = "b200",
        interpolation: Optional[Literal["Nearest", "Linear", "Smoothstep"]] = None,
    ) -> None:
        super().__init__(in_dim=3)
        if implementation not in IMPLEMENTATIONS:
            raise ValueError(f"implementation must be one of {IMPLEMENTATIONS}, got {implementation!r}")
        assert interpolation is None or interpolation == "Linear", (
            f"interpolation '{interpolation}' is not supported for the b200 encoding backend")  # encodings.py:316-319
        if features_per_level not in (1, 2, 4, 8):
            raise ValueError("features_per_level must be 1, 2, 4 or 8")
        self.num_levels = num_levels
        self.features_per_level = features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2 ** log2_hashmap_size

        # same op sequence as encodings.py:281-284 so the float32 rounding of the scales is identical
        levels = torch.arange(num_levels)
        growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        self.scalings = torch.floor(min_res * growth_factor ** levels)
        self._scalings_host = tuple(float(s) for s in self.scalings.tolist())
        self.hash_offset = levels * self.hash_table_size
        self.tcnn_encoding = None

        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1  # encodings.py:311-314
        table *= hash_init_scale
        self.hash_table = nn.Parameter(table)

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def forward(self, in_tensor: Tensor) -> Tensor:
        assert in_tensor.shape[-1] == 3  # encodings.py:346
        return ops.hash_encode(in_tensor, self.hash_table, self._scalings_host, self.log2_hashmap_size)

    def corner_indices(self, in_tensor: Tensor):
        """Parity probe: the 8 table rows per (point, level) in the reference corner order, plus offsets."""
        return ops.hash_indices(in_tensor, self._scalings_host, self.log2_hashmap_size)


class SHEncoding(Encoding):
    """Spherical harmonic encoding (encodings.py:679-719), degree 4 only (the one PreSight uses)."""

    def __init__(self, levels: int = 4, implementation: str = "b200") -> None:
        super().__init__(in_dim=3)
        if levels <= 0 or levels > 4:
            raise ValueError(f"Spherical harmonic encoding only supports 1 to 4 levels, requested {levels}")
        if levels != 4:
            raise NotImplementedError("the b200 SH kernel implements levels=4 (PreSight's setting)")
        self.levels = levels

    def get_out_dim(self) -> int:
        return self.levels ** 2

    @torch.no_grad()
    def forward(self, in_tensor: Tensor) -> Tensor:
        """`in_tensor` is the (d+1)/2-mapped direction, as the reference passes it (base_field.py:136-142)."""
        return ops.sh4(in_tensor, mapped=True)

    @torch.no_grad()
    def forward_raw(self, directions: Tensor) -> Tensor:
        """Raw unit directions; the kernel applies the (d+1)/2 mapping itself."""
        return ops.sh4(directions)
