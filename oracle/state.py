"""Build oracle parameter containers from reference-style state dicts (test infrastructure only).

Key names follow the module nesting of the reference (SURVEY §8b "State-dict compatibility"):
  iNGPField:            aabb, mlp_base_grid.hash_table, mlp_base_mlp.layers.{i}.{weight,bias},
                        semantic_head.layers.{i}.*, rgb_head.layers.{i}.*      (NGP:96-161)
  PropNetDensityField:  aabb, encoding.hash_table, mlp_base.1.layers.{i}.* | linear.*   (PROP:66-98)
  SkyField:             rgb_head.layers.{i}.*, semantic_head.layers.{i}.*      (SKY:75-93)
  *MS wrappers:         centroids, fields.{j}.<above>                          (NGPM:76-78)
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional

import torch
from torch import Tensor

from .nerf_oracle import HashGrid, Mlp, NgpField, PropField, SkyField, hash_scalings


def _t(v) -> Tensor:
    return v if isinstance(v, Tensor) else torch.as_tensor(v)


def mlp_from_state(sd: Mapping[str, Tensor], prefix: str, out_act: Optional[str] = None,
                   requires_grad: bool = False) -> Mlp:
    ws: List[Tensor] = []
    bs: List[Tensor] = []
    i = 0
    while f"{prefix}layers.{i}.weight" in sd:
        ws.append(_t(sd[f"{prefix}layers.{i}.weight"]).clone().float().requires_grad_(requires_grad))
        bs.append(_t(sd[f"{prefix}layers.{i}.bias"]).clone().float().requires_grad_(requires_grad))
        i += 1
    assert ws, f"no layers under {prefix}"
    return Mlp(ws, bs, out_act)


def grid_from_state(sd: Mapping[str, Tensor], key: str, num_levels: int, min_res: int, max_res: int,
                    log2_T: int, requires_grad: bool = False) -> HashGrid:
    table = _t(sd[key]).clone().float().requires_grad_(requires_grad)
    return HashGrid(table, hash_scalings(num_levels, min_res, max_res), log2_T)


def ngp_from_state(sd: Mapping[str, Tensor], prefix: str, meta: Dict, requires_grad: bool = False) -> NgpField:
    grid = grid_from_state(sd, prefix + "mlp_base_grid.hash_table", meta["num_levels"], meta["base_res"],
                           meta["max_res"], meta["log2_hashmap_size"], requires_grad)
    sem_dim = int(meta.get("semantic_dim", 0)) if meta.get("use_semantics", False) else 0
    return NgpField(
        aabb=_t(sd[prefix + "aabb"]).float(),
        grid=grid,
        base=mlp_from_state(sd, prefix + "mlp_base_mlp.", None, requires_grad),
        rgb=mlp_from_state(sd, prefix + "rgb_head.", "sigmoid", requires_grad),
        sem=mlp_from_state(sd, prefix + "semantic_head.", None, requires_grad) if sem_dim else None,
        geo_dim=int(meta.get("geo_feat_dim", 15)),
        sem_dim=sem_dim,
        contract=bool(meta.get("contract", True)),
    )


def prop_from_state(sd: Mapping[str, Tensor], prefix: str, meta: Dict, requires_grad: bool = False) -> PropField:
    grid = grid_from_state(sd, prefix + "encoding.hash_table", meta["num_levels"], meta["base_res"],
                           meta["max_res"], meta["log2_hashmap_size"], requires_grad)
    if meta.get("use_linear", False):
        net = Mlp([_t(sd[prefix + "linear.weight"]).clone().float().requires_grad_(requires_grad)],
                  [_t(sd[prefix + "linear.bias"]).clone().float().requires_grad_(requires_grad)], None)
    else:
        net = mlp_from_state(sd, prefix + "mlp_base.1.", None, requires_grad)
    return PropField(_t(sd[prefix + "aabb"]).float(), grid, net, bool(meta.get("contract", True)))


def sky_from_state(sd: Mapping[str, Tensor], prefix: str, use_semantics: bool, requires_grad: bool = False) -> SkyField:
    return SkyField(
        rgb=mlp_from_state(sd, prefix + "rgb_head.", "sigmoid", requires_grad),
        sem=mlp_from_state(sd, prefix + "semantic_head.", None, requires_grad) if use_semantics else None,
    )
