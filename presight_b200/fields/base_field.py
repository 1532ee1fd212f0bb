"""Field base class (reference: nerfstudio/fields/base_field.py:40-142)."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor, nn


def get_normalized_directions(directions: Tensor) -> Tensor:
    """SH encoding input range [0,1] (base_field.py:136-142)."""
    return (directions + 1.0) / 2.0


class Field(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self._aabb_cache = None

    def aabb_host(self) -> Sequence[float]:
        """The aabb buffer as 6 host floats, cached until the buffer is modified (no per-step device sync)."""
        aabb: Tensor = self.aabb
        key = (aabb.data_ptr(), aabb._version, aabb.device)
        if self._aabb_cache is None or self._aabb_cache[0] != key:
            self._aabb_cache = (key, [float(v) for v in aabb.detach().flatten().cpu().tolist()])
        return self._aabb_cache[1]
