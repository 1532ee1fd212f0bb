"""The loss stack of the hot path (reference: model_components/losses.py and model_components/PreSight/losses.py;
SURVEY §8(f)-1, the first "next" row): the terms that sit between compositing-forward and compositing-backward of every
training step.  The proposal losses (z-anti-aliased and plain), the distortion loss and the rgb / sky / semantic terms
each run as ONE kernel producing the loss and its gradient (`ps_zaa_interlevel_loss`, `ps_interlevel_loss`,
`ps_distortion_loss`, `ps_render_losses`); they take CUDA tensors only — there is no torch fallback for them (the CPU
restatements live in `oracle/`).  The depth-supervision terms (expected LiDAR / mono depth, line of sight) are one more
kernel, `ps_depth_losses`."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


def interlevel_loss(weights_list: List[Tensor], sp_bins_list: List[Tensor]) -> Tensor:
    """Proposal loss of mip-NeRF 360 (losses.py:48-126: outer / lossfun_outer / interlevel_loss); sp_bins_list holds the
    spacing-domain bin edges.  One kernel per proposal level: loss and d loss / d proposal weights (csrc/losses.cu)."""
    from . import ops
    c = sp_bins_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0.0
    for sdist, weights in zip(sp_bins_list[:-1], weights_list[:-1]):
        loss = loss + ops.interlevel_loss_level(c, w, sdist, weights)     # [N,Sp,1]: no slicing node in between
    return loss


def z_anti_aliasing_interlevel_loss(weights_list: List[Tensor], sp_bins_list: List[Tensor],
                                    pulse_width=(0.03, 0.003)) -> Tensor:
    """zip-NeRF proposal loss, the reference's default (`enable_z_anti_aliasing`, nerfacto_nusc_ms.py:129,293-295;
    model_components/PreSight/losses.py:166-206).  One kernel per proposal level (`ps_zaa_interlevel_loss`: loss and
    d loss / d proposal weights), CUDA tensors only."""
    from . import ops
    c = sp_bins_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0.0
    for i, (sdist, weights) in enumerate(zip(sp_bins_list[:-1], weights_list[:-1])):
        loss = loss + ops.zaa_interlevel_loss_level(c, w, sdist, weights, pulse_width[i])
    return loss


def distortion_loss(weights_list: List[Tensor], sp_bins_list: List[Tensor]) -> Tensor:
    """losses.py:130-149 (lossfun_distortion / distortion_loss): distortion of the final level's weights along the
    spacing-domain bins; one kernel producing the loss and d loss / d weights (`ps_distortion_loss`)."""
    from . import ops
    return ops.distortion_loss(sp_bins_list[-1].detach(), weights_list[-1])


# ---- depth supervision (PreSight/losses.py:24-103): one kernel for both terms and both gradients (`ps_depth_losses`,
# csrc/depth_loss.cu); CUDA tensors only, like the rest of the loss stack.
def depth_supervision_losses(weights: Tensor, expected_depth: Tensor, target_depth_m: Tensor, sky_mask: Optional[Tensor],
                             pose_scale: float, sigma: float, upper_bound: float, inverse: bool = False,
                             eu_bins: Optional[Tensor] = None, steps_m: Optional[Tensor] = None) -> Tensor:
    """-> [2] = (expected-depth loss, line-of-sight loss), each the reference's mean over the rays with
    1 m < target < upper_bound (and sky == 0 when a sky mask is given), before their multipliers.

    expected-depth term: `expected_depth_loss` (:67-81, LiDAR: sky_mask None) / `expected_monodepth_loss` (:83-103, with the
    sky mask, `inverse` = compare 1 / (d + 5)); line-of-sight term: `line_of_sight_loss` (:28-65) on the final level's
    weights with sample mid-points from `eu_bins` (scene units, divided by `pose_scale` like nerfacto_nusc_ms.py:584-586)
    or given in metres as `steps_m`."""
    from . import ops
    return ops.depth_losses(weights, expected_depth, target_depth_m, sky_mask, pose_scale, sigma, upper_bound, inverse,
                            eu_bins=eu_bins, steps_m=steps_m)


def expected_monodepth_loss(termination_depth: Tensor, predicted_depth: Tensor, sky_mask: Tensor,
                            upper_bound: float = 50.0, inverse: bool = False) -> Tensor:
    """PreSight/losses.py:83-103 (depths in metres)."""
    from . import ops
    return ops.depth_losses(None, predicted_depth, termination_depth, sky_mask, 1.0, 1.0, upper_bound, inverse)[0]


def expected_depth_loss(termination_depth: Tensor, predicted_depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    """PreSight/losses.py:67-81 (depths in metres)."""
    from . import ops
    return ops.depth_losses(None, predicted_depth, termination_depth, None, 1.0, 1.0, upper_bound, False)[0]


def line_of_sight_loss(weights: Tensor, termination_depth: Tensor, steps: Tensor, sigma: float,
                       sky_mask: Optional[Tensor] = None, upper_bound: float = 75.0) -> Tensor:
    """PreSight/losses.py:28-65.  weights [N,S,1], termination_depth [N,1], steps [N,S,1] (sample mid-points, metres)."""
    from . import ops
    return ops.depth_losses(weights, None, termination_depth, sky_mask, 1.0, sigma, upper_bound, False, steps_m=steps)[1]


def render_losses(outputs, batch, use_sky: bool = True, use_semantics: bool = True) -> Tensor:
    """[rgb_loss, sky_loss, semantic_loss] of get_loss_dict (nerfacto_nusc_ms.py:558-576, 641-645) before their
    multipliers: ONE kernel producing the three means and their gradients (`ps_render_losses`, CUDA tensors only).
    batch keys: "rgb" [N,3], "sky" [N,1] (1 = sky), "features" [N,C]."""
    from . import ops
    rgb, acc = outputs["rgb"], outputs["accumulation"].view(-1, 1)
    sem = outputs.get("semantics") if use_semantics else None
    return ops.render_losses(rgb, batch["rgb"], acc if use_sky else None, batch["sky"].view(-1, 1) if use_sky else None,
                             sem, batch["features"] if sem is not None else None)
