"""Scene contraction (reference: nerfstudio/field_components/spatial_distortions.py:30-90).

The contraction itself is fused into the position prologue kernel (`ps_normalize_positions`); this class only
carries the configuration, like the reference's module carries `order`.
"""
from typing import Optional, Union

from torch import nn


class SpatialDistortion(nn.Module):
    pass


class SceneContraction(SpatialDistortion):
    def __init__(self, order: Optional[Union[float, int]] = None) -> None:
        super().__init__()
        if order != float("inf"):
            raise NotImplementedError("the b200 position kernel implements the L-inf contraction PreSight uses")
        self.order = order
