"""The loss stack of the hot path (reference: model_components/losses.py and model_components/PreSight/losses.py;
SURVEY §8(f)-1, the first "next" row): the terms that sit between compositing-forward and compositing-backward of every
training step.  The proposal losses (z-anti-aliased and plain), the distortion loss and the rgb / sky / semantic terms
each run as ONE kernel producing the loss and its gradient (`ps_zaa_interlevel_loss`, `ps_interlevel_loss`,
`ps_distortion_loss`, `ps_render_losses`); they take CUDA tensors only — there is no torch fallback for them (the CPU
restatements live in `oracle/`).  The depth-supervision terms at the end are the rows not yet written as kernels and are
plain torch expressions."""
from __future__ import annotations

from typing import List

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


# ---- depth supervision (PreSight/losses.py:25-103).  These three are still plain torch expressions (SURVEY 8f-1 lists
# them as the remaining rows of the loss stack); they are reached only when the batch carries a "depth" target.
URF_SIGMA_SCALE_FACTOR = 3.0


def normalize_depth(depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    return torch.clip(depth / upper_bound, 0.0, 1.0)


def expected_monodepth_loss(termination_depth: Tensor, predicted_depth: Tensor, sky_mask: Tensor,
                            upper_bound: float = 50.0, inverse: bool = False) -> Tensor:
    """PreSight/losses.py:83-103."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound) & (sky_mask == 0.0)
    if inverse:
        termination_depth, predicted_depth = 1 / (termination_depth + 5), 1 / (predicted_depth + 5)
    else:
        termination_depth = normalize_depth(termination_depth, upper_bound)
        predicted_depth = normalize_depth(predicted_depth, upper_bound)
    return torch.mean(((termination_depth - predicted_depth) ** 2)[depth_mask])


def expected_depth_loss(termination_depth: Tensor, predicted_depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    """PreSight/losses.py:67-81."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound)
    diff = normalize_depth(termination_depth, upper_bound) - normalize_depth(predicted_depth, upper_bound)
    return torch.mean((diff ** 2)[depth_mask])


def line_of_sight_loss(weights: Tensor, termination_depth: Tensor, steps: Tensor, sigma: float,
                       sky_mask: Tensor = None, upper_bound: float = 75.0) -> Tensor:
    """PreSight/losses.py:28-65 (Urban Radiance Fields): weights [N,S,1], termination_depth [N,1], steps [N,S,1]."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound)
    if sky_mask is not None:
        depth_mask = depth_mask & (sky_mask == 0.0)
    steps = steps.detach()
    td = termination_depth[:, None]
    target = torch.distributions.normal.Normal(0.0, sigma / URF_SIGMA_SCALE_FACTOR)
    near_mask = torch.logical_and(steps <= td + sigma, steps >= td - sigma)
    near = (near_mask * (weights - torch.exp(target.log_prob(steps - td))) ** 2).sum(-2)
    empty = ((steps < td - sigma) * weights ** 2).sum(-2)
    return torch.mean((near + empty)[depth_mask])


def render_losses(outputs, batch, use_sky: bool = True, use_semantics: bool = True) -> Tensor:
    """[rgb_loss, sky_loss, semantic_loss] of get_loss_dict (nerfacto_nusc_ms.py:558-576, 641-645) before their
    multipliers: ONE kernel producing the three means and their gradients (`ps_render_losses`, CUDA tensors only).
    batch keys: "rgb" [N,3], "sky" [N,1] (1 = sky), "features" [N,C]."""
    from . import ops
    rgb, acc = outputs["rgb"], outputs["accumulation"].view(-1, 1)
    sem = outputs.get("semantics") if use_semantics else None
    return ops.render_losses(rgb, batch["rgb"], acc if use_sky else None, batch["sky"].view(-1, 1) if use_sky else None,
                             sem, batch["features"] if sem is not None else None)
