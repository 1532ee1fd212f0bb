"""The loss stack of the hot path (reference: model_components/losses.py and model_components/PreSight/losses.py;
SURVEY §8(f)-1, the first "next" row): the terms that sit between compositing-forward and compositing-backward of every
training step.  The proposal losses (z-anti-aliased and plain), the distortion loss and the rgb / sky / semantic terms
each run as ONE kernel producing the loss and its gradient (`ps_zaa_interlevel_loss`, `ps_interlevel_loss`,
`ps_distortion_loss`, `ps_render_losses`); they take CUDA tensors only — there is no torch fallback for them (the CPU
restatements live in `oracle/`).  The depth-supervision terms at the end are the rows not yet written as kernels and are
plain torch expressions."""
from __future__ import annotations

import math
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


# ---- depth supervision (PreSight/losses.py:25-103).  Not kernels yet (SURVEY 8f-1 lists them as the remaining rows of the
# loss stack): per-ray torch expressions, reached only when the batch carries a "depth" target.
def _supervised_rays(target_m: Tensor, limit_m: float, sky_mask: Optional[Tensor]) -> Tensor:
    """Rays whose target depth is usable: strictly between 1 m and the upper bound and, if a sky mask is given, not sky."""
    ok = (target_m > 1.0) & (target_m < limit_m)
    return ok if sky_mask is None else ok & (sky_mask == 0.0)


def normalize_depth(depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    return (depth / upper_bound).clip(0.0, 1.0)


def expected_monodepth_loss(termination_depth: Tensor, predicted_depth: Tensor, sky_mask: Tensor,
                            upper_bound: float = 50.0, inverse: bool = False) -> Tensor:
    """Mono-depth supervision of the rendered expected depth (:83-103): squared error of the depths mapped to [0, 1]
    (or to 1 / (d + 5) when `inverse`), averaged over the supervised, non-sky rays."""
    rays = _supervised_rays(termination_depth, upper_bound, sky_mask)
    if inverse:
        err = 1 / (termination_depth + 5) - 1 / (predicted_depth + 5)
    else:
        err = normalize_depth(termination_depth, upper_bound) - normalize_depth(predicted_depth, upper_bound)
    return err.square()[rays].mean()


def expected_depth_loss(termination_depth: Tensor, predicted_depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    """LiDAR supervision of the rendered expected depth (:67-81): as above, without the sky mask."""
    rays = _supervised_rays(termination_depth, upper_bound, None)
    err = normalize_depth(termination_depth, upper_bound) - normalize_depth(predicted_depth, upper_bound)
    return err.square()[rays].mean()


def line_of_sight_loss(weights: Tensor, termination_depth: Tensor, steps: Tensor, sigma: float,
                       sky_mask: Optional[Tensor] = None, upper_bound: float = 75.0) -> Tensor:
    """Line-of-sight loss of Urban Radiance Fields (:28-65).  weights [N,S,1], termination_depth [N,1], steps [N,S,1]
    (sample mid-points, metres).  Within +-sigma of the target the weights should follow N(0, sigma / 3) evaluated at
    the signed distance; every sample more than sigma in front of the target should carry no weight."""
    rays = _supervised_rays(termination_depth, upper_bound, sky_mask)
    at, target = steps.detach(), termination_depth[:, None]                # [N,S,1], [N,1,1]
    std = sigma / 3.0
    gauss = torch.exp(-((at - target) ** 2) / (2 * std ** 2) - math.log(std) - math.log(math.sqrt(2 * math.pi)))
    band = (at <= target + sigma) & (at >= target - sigma)                 # the reference's comparisons, as written
    per_ray = (band * (weights - gauss).square()).sum(-2) + ((at < target - sigma) * weights.square()).sum(-2)
    return per_ray[rays].mean()


def render_losses(outputs, batch, use_sky: bool = True, use_semantics: bool = True) -> Tensor:
    """[rgb_loss, sky_loss, semantic_loss] of get_loss_dict (nerfacto_nusc_ms.py:558-576, 641-645) before their
    multipliers: ONE kernel producing the three means and their gradients (`ps_render_losses`, CUDA tensors only).
    batch keys: "rgb" [N,3], "sky" [N,1] (1 = sky), "features" [N,C]."""
    from . import ops
    rgb, acc = outputs["rgb"], outputs["accumulation"].view(-1, 1)
    sem = outputs.get("semantics") if use_semantics else None
    return ops.render_losses(rgb, batch["rgb"], acc if use_sky else None, batch["sky"].view(-1, 1) if use_sky else None,
                             sem, batch["features"] if sem is not None else None)
