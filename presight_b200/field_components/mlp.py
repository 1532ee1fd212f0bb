"""Drop-in MLP (reference: nerfstudio/field_components/mlp.py:77-179).

Same constructor, same `layers: ModuleList[nn.Linear]` (so checkpoints load), but forward/backward are ONE
fused tensor-core kernel each (`ps_mlp_fwd` / `ps_mlp_bwd`): hidden activations never reach HBM.
`implementation="b200"` uses bf16 MMA with fp32 accumulate; `"b200+fp32"` uses error-compensated 3xTF32.
"""
from __future__ import annotations

from typing import Literal, Optional, Set, Tuple

import torch
from torch import Tensor, nn

from .. import ops


class MLP(nn.Module):
    def __init__(
        self,
        in_dim: int,
        num_layers: int,
        layer_width: int,
        out_dim: Optional[int] = None,
        skip_connections: Optional[Tuple[int]] = None,
        activation: Optional[nn.Module] = nn.ReLU(),
        out_activation: Optional[nn.Module] = None,
        implementation: Literal["b200", "b200+fp32"] = "b200",
    ) -> None:
        super().__init__()
        self.in_dim = in_dim
        assert self.in_dim > 0
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers = num_layers
        self.layer_width = layer_width
        self.skip_connections = skip_connections
        self._skip_connections: Set[int] = set(skip_connections) if skip_connections else set()
        if self._skip_connections:
            raise NotImplementedError("skip connections are not used by PreSight and not implemented by the b200 MLP")
        if activation is not None and not isinstance(activation, nn.ReLU):
            raise NotImplementedError("the b200 MLP implements ReLU hidden activations")
        if out_activation is None:
            self._out_act = ops.ACT_NONE
        elif isinstance(out_activation, nn.Sigmoid):
            self._out_act = ops.ACT_SIGMOID
        elif isinstance(out_activation, nn.ReLU):
            self._out_act = ops.ACT_RELU
        else:
            raise NotImplementedError(f"output activation {out_activation} not implemented by the b200 MLP")
        if implementation not in ("b200", "b200+fp32"):
            raise ValueError(f"implementation must be 'b200' or 'b200+fp32', got {implementation!r}")
        self.activation = activation
        self.out_activation = out_activation
        self.precision = ops.PREC_BF16 if implementation == "b200" else ops.PREC_TF32X3
        self.tcnn_encoding = None
        self.build_nn_modules()

    def build_nn_modules(self) -> None:
        """Same parameter layout as mlp.py:138-155."""
        layers = []
        if self.num_layers == 1:
            layers.append(nn.Linear(self.in_dim, self.out_dim))
        else:
            for i in range(self.num_layers - 1):
                layers.append(nn.Linear(self.in_dim if i == 0 else self.layer_width, self.layer_width))
            layers.append(nn.Linear(self.layer_width, self.out_dim))
        self.layers = nn.ModuleList(layers)

    def forward(self, in_tensor: Tensor) -> Tensor:
        return ops.mlp(in_tensor, [l.weight for l in self.layers], [l.bias for l in self.layers], self._out_act,
                       self.precision)
