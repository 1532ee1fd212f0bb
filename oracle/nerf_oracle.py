"""Torch-CPU restatement of the reference's PyTorch path (test infrastructure only).

Every function cites the reference lines it restates.  Shorthand for paths:
  ENC  = field_components/encodings.py        MLPF = field_components/mlp.py
  ACT  = field_components/activations.py      SD   = field_components/spatial_distortions.py
  NGP  = fields/PreSight/ingp_field.py        PROP = fields/PreSight/prop_density_field.py
  NGPM = fields/PreSight/ingp_field_ms.py     SKY  = fields/PreSight/sky_field.py
  FU   = fields/PreSight/utils.py             BF   = fields/base_field.py
  RS   = model_components/ray_samplers.py     RN   = model_components/renderers.py
  RAYS = cameras/rays.py                      MATH = utils/math.py
  MODEL= models/PreSight/nerfacto_nusc_ms.py  XP   = scripts/extract_priors.py
  LS   = model_components/losses.py           PL   = model_components/PreSight/losses.py

The code is written functionally over plain tensors (no nn.Module tree) so that the
same functions serve as (a) checker for the CUDA kernels, (b) generator-independent
restatement that is itself checked against fixtures produced by the live reference.
Gradients come from torch autograd, exactly as in the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

__all__ = [
    "HashGrid", "Mlp", "NgpField", "PropField", "SkyField", "Model", "ModelCfg",
    "hash_scalings", "hash_corner_indices", "hash_encode", "normalize_to_unit_cube",
    "mlp_forward", "trunc_exp", "sh4", "ngp_density", "ngp_heads", "ngp_forward", "ngp_semantics",
    "prop_density", "nearest_centroid", "ms_dispatch", "spaced_bins", "spacing_fns", "pdf_cdf",
    "pdf_resample", "sample_positions", "get_weights", "render_rgb", "render_accumulation",
    "render_depth_expected", "render_depth_threshold", "proposal_sample", "model_outputs",
    "model_depth", "prior_query", "sky_outputs", "PRIME_Y", "PRIME_Z", "mlp_forward_bf16_emulated",
    "loss_outer", "lossfun_outer", "interlevel_loss", "sky_blend", "rgb_loss", "sky_loss", "semantic_loss",
    "lossfun_distortion", "distortion_loss", "blur_stepfun", "sorted_interp_quad", "z_anti_aliasing_interlevel_loss",
    "normalize_depth", "expected_monodepth_loss", "expected_depth_loss", "line_of_sight_loss", "line_of_sight_sigma",
    "line_of_sight_mult", "pinhole_rays",
]

PRIME_Y = 2654435761  # ENC:336
PRIME_Z = 805459861   # ENC:336


# --------------------------------------------------------------------------------------
# parameter containers (plain tensors; key names follow the reference's state_dict)
# --------------------------------------------------------------------------------------
@dataclass
class HashGrid:
    table: Tensor        # [L*T, F] fp32                    (ENC:311-314)
    scalings: Tensor     # [L] fp32                         (ENC:282-284)
    log2_T: int

    @property
    def L(self) -> int:
        return int(self.scalings.numel())

    @property
    def F(self) -> int:
        return int(self.table.shape[1])


@dataclass
class Mlp:
    weights: List[Tensor]            # each [out, in]       (MLPF:138-155)
    biases: List[Tensor]
    out_act: Optional[str] = None    # None | "sigmoid"     (NGP:153-161)


@dataclass
class NgpField:
    aabb: Tensor                     # [2,3]
    grid: HashGrid
    base: Mlp                        # L*F -> 64 -> 1+geo+sem   (NGP:130-138)
    rgb: Mlp                         # 16+geo+app -> 64 -> 64 -> 3 sigmoid
    sem: Optional[Mlp] = None        # sem -> 64 -> 64 -> sem   (NGP:142-151)
    geo_dim: int = 15
    sem_dim: int = 0
    contract: bool = True


@dataclass
class PropField:
    aabb: Tensor
    grid: HashGrid
    net: Mlp                         # L*F -> hidden -> 1, or single Linear (PROP:85-98)
    contract: bool = True


@dataclass
class SkyField:
    rgb: Mlp
    sem: Optional[Mlp] = None


@dataclass
class ModelCfg:
    num_proposal_samples: Tuple[int, ...] = (128, 64)
    num_nerf_samples: int = 64
    near: float = 0.1 * 0.05
    far: float = 1000.0 * 0.05
    piecewise_thr: float = 100.0 * 0.05
    histogram_padding: float = 0.01
    single_jitter: bool = True


@dataclass
class Model:
    cfg: ModelCfg
    centroids: Tensor                       # [nf,3]
    fields: List[NgpField]                  # one per sub-field
    props: List[List[PropField]]            # [level][sub-field]
    sky: Optional[List[SkyField]] = None


# --------------------------------------------------------------------------------------
# a1-a3  multiresolution hash encoding
# --------------------------------------------------------------------------------------
def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """Per-level scale factors, float32.  Restates ENC:281-284.

    The reference raises a *numpy float64 scalar* to a *torch int64 tensor*, which torch
    evaluates in float32 — so e.g. the top level of 16→2048 is 2047, not 2048.  Use the
    same torch op so the float32 rounding is reproduced (pinned by the SURVEY §8c KATs).
    """
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth ** levels).to(torch.float32)


def _hash3(ix: Tensor, iy: Tensor, iz: Tensor, log2_T: int) -> Tensor:
    """Spatial hash of integer grid coordinates (int64 in, int64 out in [0,T)).  ENC:336-339.

    int64 multiply / xor / floor-mod by 2^k equals uint32 wrap-around arithmetic masked to
    k bits, negative coordinates included (two's complement low bits).
    """
    h = ix.to(torch.int64) ^ (iy.to(torch.int64) * PRIME_Y) ^ (iz.to(torch.int64) * PRIME_Z)
    return h & ((1 << log2_T) - 1)


# corner order of ENC:354-361: bit pattern (x is ceil?, y is ceil?, z is ceil?)
_CORNER_IS_CEIL = (
    (1, 1, 1),  # h0
    (1, 0, 1),  # h1
    (0, 0, 1),  # h2
    (0, 1, 1),  # h3
    (1, 1, 0),  # h4
    (1, 0, 0),  # h5
    (0, 0, 0),  # h6
    (0, 1, 0),  # h7
)


def hash_corner_indices(x01: Tensor, scalings: Tensor, log2_T: int) -> Tuple[Tensor, Tensor]:
    """Rows of the 8 corners per (point, level) and the fractional offsets.

    Restates ENC:346-361.  Returns (idx [P,L,8] int64 incl. the level·T offset of ENC:340,
    offset [P,L,3] fp32).  ceil/floor are taken independently: on an exact integer
    coordinate both corners alias the same entry and the offset is 0.
    """
    assert x01.shape[-1] == 3
    x = x01.reshape(-1, 1, 3).to(torch.float32)
    scaled = x * scalings.view(-1, 1)                        # [P,L,3] fp32 multiply
    up = torch.ceil(scaled).to(torch.int32)
    dn = torch.floor(scaled).to(torch.int32)
    offset = scaled - dn
    T = 1 << log2_T
    level_base = torch.arange(scalings.numel(), dtype=torch.int64) * T
    cols = []
    for cx, cy, cz in _CORNER_IS_CEIL:
        ix = up[..., 0] if cx else dn[..., 0]
        iy = up[..., 1] if cy else dn[..., 1]
        iz = up[..., 2] if cz else dn[..., 2]
        cols.append(_hash3(ix, iy, iz, log2_T) + level_base)
    return torch.stack(cols, dim=-1), offset


def hash_encode(x01: Tensor, grid: HashGrid) -> Tensor:
    """[...,3] in the unit cube -> [..., L*F] features.  Restates ENC:343-384.

    Interpolation order is the reference's: pairs along x (03,12,56,47), then y, then z,
    each as `hi*o + lo*(1-o)` in fp32.
    """
    lead = x01.shape[:-1]
    idx, off = hash_corner_indices(x01, grid.scalings, grid.log2_T)
    t = [grid.table[idx[..., k]] for k in range(8)]          # each [P,L,F]
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    f03 = t[0] * ox + t[3] * (1 - ox)
    f12 = t[1] * ox + t[2] * (1 - ox)
    f56 = t[5] * ox + t[6] * (1 - ox)
    f47 = t[4] * ox + t[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    out = f0312 * oz + f4756 * (1 - oz)
    return out.reshape(*lead, grid.L * grid.F)


# --------------------------------------------------------------------------------------
# a4-a5  position normalisation, L-inf contraction, selector
# --------------------------------------------------------------------------------------
def normalize_to_unit_cube(pos: Tensor, aabb: Tensor, contract: bool = True) -> Tuple[Tensor, Tensor]:
    """World position -> ([0,1]^3 position with masked points moved to the origin, selector).

    Restates FU:6-10, SD:66-69 (order=inf), NGP:169-177 (identical in PROP:130-138).
    Without contraction the reference uses SceneBox.get_normalized_positions
    (data/scene_box.py:57-66): (p - min) / (max - min).
    """
    lo, hi = aabb[0], aabb[1]
    if contract:
        p = (pos - lo) / (hi - lo)
        p = p * 2 - 1
        mag = torch.linalg.norm(p, ord=float("inf"), dim=-1)[..., None]
        p = torch.where(mag < 1, p, (2 - (1 / mag)) * (p / mag))
        p = (p + 2.0) / 4.0
    else:
        p = (pos - lo) / (hi - lo)
    sel = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    return p * sel[..., None], sel


# --------------------------------------------------------------------------------------
# a6-a7  MLP and truncated exp
# --------------------------------------------------------------------------------------
def mlp_forward(x: Tensor, net: Mlp) -> Tensor:
    """Linear+ReLU stack with optional output activation.  Restates MLPF:157-174."""
    n = len(net.weights)
    for i, (w, b) in enumerate(zip(net.weights, net.biases)):
        x = torch.nn.functional.linear(x, w, b)
        if i < n - 1:
            x = torch.relu(x)
    if net.out_act == "sigmoid":
        x = torch.sigmoid(x)
    elif net.out_act is not None:
        raise ValueError(net.out_act)
    return x


class _TruncExpOracle(torch.autograd.Function):
    """exp forward; backward multiplies by exp(clamp(x,-15,15)).  Restates ACT:28-41."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, -15, 15))


def trunc_exp(x: Tensor) -> Tensor:
    return _TruncExpOracle.apply(x.to(torch.float32))


# --------------------------------------------------------------------------------------
# SH degree-4 direction encoding
# --------------------------------------------------------------------------------------
def sh4(d01: Tensor) -> Tensor:
    """16 real SH components of the *already (d+1)/2-mapped* direction.  Restates MATH:27-74.

    The torch path of the reference evaluates the polynomials directly on the [0,1]-mapped
    direction (BF:136-142 then ENC:708-711); no gradient flows (ENC:708 is under no_grad).
    """
    with torch.no_grad():
        x, y, z = d01[..., 0], d01[..., 1], d01[..., 2]
        xx, yy, zz = x ** 2, y ** 2, z ** 2
        c = torch.zeros((*d01.shape[:-1], 16), dtype=torch.float32)
        c[..., 0] = 0.28209479177387814
        c[..., 1] = 0.4886025119029199 * y
        c[..., 2] = 0.4886025119029199 * z
        c[..., 3] = 0.4886025119029199 * x
        c[..., 4] = 1.0925484305920792 * x * y
        c[..., 5] = 1.0925484305920792 * y * z
        c[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
        c[..., 7] = 1.0925484305920792 * x * z
        c[..., 8] = 0.5462742152960396 * (xx - yy)
        c[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
        c[..., 10] = 2.890611442640554 * x * y * z
        c[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
        c[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
        c[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
        c[..., 14] = 1.445305721320277 * z * (xx - yy)
        c[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return c


# --------------------------------------------------------------------------------------
# a8-a9  fields
# --------------------------------------------------------------------------------------
def ngp_density(f: NgpField, pos: Tensor) -> Tuple[Tensor, Tensor]:
    """positions [...,3] -> (density [...,1], embedding [..., geo+sem]).  Restates NGP:168-191."""
    p, sel = normalize_to_unit_cube(pos, f.aabb, f.contract)
    h = mlp_forward(hash_encode(p.reshape(-1, 3), f.grid), f.base).view(*p.shape[:-1], -1)
    raw, emb = torch.split(h, [1, f.geo_dim + f.sem_dim], dim=-1)
    return trunc_exp(raw) * sel[..., None], emb


def ngp_heads(f: NgpField, dirs: Tensor, emb: Tensor, app: Optional[Tensor]) -> Dict[str, Tensor]:
    """Colour (+ semantics) heads.  Restates NGP:193-237."""
    out: Dict[str, Tensor] = {}
    shape = dirs.shape[:-1]
    if f.sem is not None:
        emb, sem_in = torch.split(emb, [f.geo_dim, f.sem_dim], dim=-1)
        out["semantics"] = mlp_forward(sem_in.reshape(-1, f.sem_dim), f.sem).view(*shape, -1)
    d = sh4(((dirs + 1.0) / 2.0).reshape(-1, 3))
    parts = [d, emb.reshape(-1, f.geo_dim)]
    if app is not None:
        parts.append(app.reshape(-1, app.shape[-1]))
    out["rgb"] = mlp_forward(torch.cat(parts, dim=-1), f.rgb).view(*shape, 3)
    return out


def ngp_forward(f: NgpField, pos: Tensor, dirs: Tensor, app: Optional[Tensor]) -> Dict[str, Tensor]:
    """Restates NGP:239-251."""
    density, emb = ngp_density(f, pos)
    out = ngp_heads(f, dirs, emb, app)
    out["density"] = density
    return out


def ngp_semantics(f: NgpField, pos: Tensor) -> Tensor:
    """Restates NGP:253-267 (re-runs hash + base MLP)."""
    _, emb = ngp_density(f, pos)
    _, sem_in = torch.split(emb, [f.geo_dim, f.sem_dim], dim=-1)
    return mlp_forward(sem_in.reshape(-1, f.sem_dim), f.sem).view(*pos.shape[:-1], -1)


def prop_density(f: PropField, pos: Tensor) -> Tensor:
    """positions [...,3] -> density [...,1].  Restates PROP:129-153."""
    p, sel = normalize_to_unit_cube(pos, f.aabb, f.contract)
    raw = mlp_forward(hash_encode(p.reshape(-1, 3), f.grid), f.net).view(*p.shape[:-1], -1)
    return trunc_exp(raw) * sel[..., None]


def sky_outputs(f: SkyField, dirs: Tensor, app: Optional[Tensor]) -> Dict[str, Tensor]:
    """Per-ray sky colour / semantics.  Restates SKY:95-111."""
    d = sh4((dirs + 1.0) / 2.0)
    out = {"rgb": mlp_forward(torch.cat([d, app], dim=-1) if app is not None else d, f.rgb)}
    if f.sem is not None:
        out["semantics"] = mlp_forward(d, f.sem)
    return out


# --------------------------------------------------------------------------------------
# a10  nearest-centroid routing over sub-fields
# --------------------------------------------------------------------------------------
def nearest_centroid(points: Tensor, centroids: Tensor) -> Tensor:
    """argmin_j ||p - c_j||.  Restates NGPM:97 (torch.cdist(...).argmin)."""
    return torch.cdist(points, centroids).argmin(dim=1)


def ms_dispatch(points: Tensor, centroids: Tensor, per_field, n_fields: int, extras: Sequence[Optional[Tensor]] = ()):
    """Route flat points to sub-fields, run `per_field(i, pts, *extras_i)` (returns a dict of
    [n_i, c] tensors) and scatter back.  Restates the loop of NGPM:99-126 / 133-153 /
    fields/PreSight/prop_density_field_ms.py:90-102.
    """
    which = nearest_centroid(points, centroids)
    out: Dict[str, Tensor] = {}
    for i in range(n_fields):
        m = which == i
        if not bool(m.any()):
            continue
        sub = per_field(i, points[m], *[None if e is None else e[m] for e in extras])
        for k, v in sub.items():
            if k not in out:
                out[k] = torch.empty(points.shape[0], v.shape[-1], dtype=v.dtype)
            out[k][m] = v          # in-place masked write, differentiable w.r.t. v
    return out


# --------------------------------------------------------------------------------------
# a11-a14  samplers and sample geometry
# --------------------------------------------------------------------------------------
def spacing_fns(thr: float):
    """PreSight's piecewise spacing and its inverse.  Restates MODEL:312-317."""
    def fn(x):
        return torch.where(x < thr, x / (2 * thr), 1 - 1 / (2 * x / thr))

    def inv(x):
        return torch.where(x < 0.5, x * (2 * thr), thr / (2 - 2 * x))

    return fn, inv


def spaced_bins(nears: Tensor, fars: Tensor, num_samples: int, thr: float,
                t_rand: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Initial sampler.  nears/fars [N,1]; t_rand [N,1] (single jitter) or [N,S+1] or None (eval).

    Returns (spacing bins [N,S+1] in [0,1], euclidean bins [N,S+1]).  Restates RS:98-128.
    """
    fn, inv = spacing_fns(thr)
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, ...]
    if t_rand is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * t_rand
    else:
        bins = bins.expand(nears.shape[0], -1)
    s_near, s_far = fn(nears), fn(fars)
    eu = inv(bins * s_far + (1 - bins) * s_near)
    return bins.expand(nears.shape[0], -1), eu


def pdf_cdf(weights: Tensor, padding: float, eps: float) -> Tensor:
    """weights [N,S] -> cdf [N,S+1].  Restates RS:305-315."""
    w = weights + padding
    wsum = torch.sum(w, dim=-1, keepdim=True)
    pad = torch.relu(eps - wsum)
    w = w + pad / w.shape[-1]
    wsum = wsum + pad
    pdf = w / wsum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def pdf_u(num_rays: int, num_samples: int, rand: Optional[Tensor]) -> Tensor:
    """Sample positions in CDF space.  rand: [N,1] / [N,S+1] uniform jitter (train) or None (eval).
    Restates RS:317-331."""
    nb = num_samples + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
    if rand is not None:
        u = u.expand(num_rays, nb) + rand / nb
    else:
        u = (u + 1.0 / (2 * nb)).expand(num_rays, nb)
    return u.contiguous()


def pdf_resample(weights: Tensor, existing_bins: Tensor, num_samples: int, rand: Optional[Tensor],
                 padding: float = 0.01, eps: float = 1e-5) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Inverse-CDF resampling (include_original=False).  Restates RS:305-360.

    weights [N,S_in], existing_bins [N,S_in+1] (spacing domain).
    Returns (bins [N,S_out+1] detached, inds [N,S_out+1] int64, cdf, u).
    """
    cdf = pdf_cdf(weights, padding, eps)
    u = pdf_u(weights.shape[0], num_samples, rand)
    inds = torch.searchsorted(cdf, u, side="right")
    hi_cap = existing_bins.shape[-1] - 1
    below = torch.clamp(inds - 1, 0, hi_cap)
    above = torch.clamp(inds, 0, hi_cap)
    c0, b0 = torch.gather(cdf, -1, below), torch.gather(existing_bins, -1, below)
    c1, b1 = torch.gather(cdf, -1, above), torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    bins = (b0 + t * (b1 - b0)).detach()
    return bins, inds, cdf, u


def sample_positions(origins: Tensor, dirs: Tensor, eu_bins: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Euclidean bin edges [N,S+1] -> (positions [N,S,3], starts, ends, deltas [N,S,1]).
    Restates RAYS:251-295 and RAYS:49-58 (pos = o + d*(start+end)/2)."""
    starts, ends = eu_bins[..., :-1, None], eu_bins[..., 1:, None]
    pos = origins[:, None, :] + dirs[:, None, :] * (starts + ends) / 2
    return pos, starts, ends, ends - starts


# --------------------------------------------------------------------------------------
# 8f-1  loss stack: proposal (interlevel) loss.  LS = model_components/losses.py
# --------------------------------------------------------------------------------------
LOSS_EPS = 1.0e-7   # LS:38 (torch.finfo-free constant EPS)


def loss_outer(t0_starts: Tensor, t0_ends: Tensor, t1_starts: Tensor, t1_ends: Tensor, y1: Tensor) -> Tensor:
    """Upper envelope of the step function (t1, y1) on the intervals t0.  Restates LS:47-77."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    return torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1) - torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)


def lossfun_outer(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor) -> Tensor:
    """Restates LS:80-97."""
    w_outer = loss_outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + LOSS_EPS)


def interlevel_loss(weights_list: Sequence[Tensor], sp_bins_list: Sequence[Tensor]) -> Tensor:
    """Proposal loss (LS:108-126): weights_list [N,S_k,1] per level, sp_bins_list [N,S_k+1] spacing-domain bin edges
    (= ray_samples_to_sdist, LS:100-105); the last level is the detached target."""
    c = sp_bins_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0.0
    for sdist, weights in zip(sp_bins_list[:-1], weights_list[:-1]):
        loss = loss + torch.mean(lossfun_outer(c, w, sdist, weights[..., 0]))
    return loss


def blur_stepfun(x: Tensor, y: Tensor, r: float) -> Tuple[Tensor, Tensor]:
    """Step function (x [.., S+1] edges, y [.., S] heights) convolved with a box of half-width r -> piecewise-linear
    (xr [.., 2S+2], yr [.., 2S+2]).  Restates model_components/PreSight/losses.py:127-139."""
    xr, xr_idx = torch.sort(torch.cat([x - r, x + r], dim=-1))
    y1 = (torch.cat([y, torch.zeros_like(y[..., :1])], dim=-1)
          - torch.cat([torch.zeros_like(y[..., :1]), y], dim=-1)) / (2 * r)
    y2 = torch.cat([y1, -y1], dim=-1).take_along_dim(xr_idx[..., :-1], dim=-1)
    yr = torch.cumsum((xr[..., 1:] - xr[..., :-1]) * torch.cumsum(y2, dim=-1), dim=-1).clamp_min(0)
    yr = torch.cat([torch.zeros_like(yr[..., :1]), yr], dim=-1)
    return xr, yr


def sorted_interp_quad(x: Tensor, xp: Tensor, fpdf: Tensor, fcdf: Tensor) -> Tensor:
    """Integral of the piecewise-linear density (knots xp, values fpdf, running integral fcdf) evaluated at the sorted
    query points x.  Restates PreSight/losses.py:141-164."""
    mask = x[..., None, :] >= xp[..., :, None]

    def find_interval(v, return_idx=False):
        v0, i0 = torch.max(torch.where(mask, v[..., None], v[..., :1, None]), -2)
        v1, i1 = torch.min(torch.where(~mask, v[..., None], v[..., -1:, None]), -2)
        return (v0, v1, i0, i1) if return_idx else (v0, v1)

    fcdf0, fcdf1, i0, i1 = find_interval(fcdf, return_idx=True)
    fpdf0 = fpdf.take_along_dim(i0, dim=-1)
    fpdf1 = fpdf.take_along_dim(i1, dim=-1)
    xp0, xp1 = find_interval(xp)
    offset = torch.clip(torch.nan_to_num((x - xp0) / (xp1 - xp0), 0), 0, 1)
    return fcdf0 + (x - xp0) * (fpdf0 + fpdf1 * offset + fpdf0 * (1 - offset)) / 2


def z_anti_aliasing_interlevel_loss(weights_list: Sequence[Tensor], sp_bins_list: Sequence[Tensor],
                                    pulse_width: Sequence[float]) -> Tensor:
    """zip-NeRF proposal loss, the reference's default (`enable_z_anti_aliasing`, MODEL:129,293-295).  Restates
    PreSight/losses.py:166-206: the final level's histogram, blurred with a per-level pulse width, is integrated over
    each proposal level's bins and the proposal weights are pushed up to that envelope."""
    c = sp_bins_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    w_normalized = w / (c[..., 1:] - c[..., :-1])
    loss = 0.0
    for i, (cp, weights) in enumerate(zip(sp_bins_list[:-1], weights_list[:-1])):
        ci, wi = blur_stepfun(c, w_normalized, pulse_width[i])
        area = 0.5 * (wi[..., 1:] + wi[..., :-1]) * (ci[..., 1:] - ci[..., :-1])
        cdfs = torch.cat([torch.zeros_like(area[..., :1]), torch.cumsum(area, dim=-1)], dim=-1)
        wp = weights[..., 0]
        cdf_interp = sorted_interp_quad(cp, ci, wi, cdfs)
        w_s = torch.diff(cdf_interp, dim=-1)
        loss = loss + ((w_s - wp).clamp_min(0) ** 2 / (wp + 1e-5)).mean()
    return loss


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    """LS:130-143: sum_ij w_i w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3 per ray, u = bin mid-points."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(weights_list: Sequence[Tensor], sp_bins_list: Sequence[Tensor]) -> Tensor:
    """LS:145-149 (the final level's weights are NOT detached here: the gradient reaches the field)."""
    return torch.mean(lossfun_distortion(sp_bins_list[-1], weights_list[-1][..., 0]))


URF_SIGMA_SCALE_FACTOR = 3.0   # PL:22


def normalize_depth(depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    """PL:25-26."""
    return torch.clip(depth / upper_bound, 0.0, 1.0)


def expected_monodepth_loss(termination_depth: Tensor, predicted_depth: Tensor, sky_mask: Tensor,
                            upper_bound: float = 50.0, inverse: bool = False) -> Tensor:
    """Monocular-depth supervision of the expected depth (model_components/PreSight/losses.py:83-103): MSE of the
    normalised (or inverse) depths over rays with 1 < depth < upper_bound that are not sky."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound) & (sky_mask == 0.0)
    if inverse:
        termination_depth = 1 / (termination_depth + 5)
        predicted_depth = 1 / (predicted_depth + 5)
    else:
        termination_depth = normalize_depth(termination_depth, upper_bound=upper_bound)
        predicted_depth = normalize_depth(predicted_depth, upper_bound=upper_bound)
    return torch.mean(((termination_depth - predicted_depth) ** 2)[depth_mask])


def expected_depth_loss(termination_depth: Tensor, predicted_depth: Tensor, upper_bound: float = 75.0) -> Tensor:
    """LiDAR variant (PL:67-81): no sky mask."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound)
    t = normalize_depth(termination_depth, upper_bound=upper_bound)
    p = normalize_depth(predicted_depth, upper_bound=upper_bound)
    return torch.mean(((t - p) ** 2)[depth_mask])


def line_of_sight_loss(weights: Tensor, termination_depth: Tensor, steps: Tensor, sigma: float,
                       sky_mask: Optional[Tensor] = None, upper_bound: float = 75.0) -> Tensor:
    """Urban-Radiance-Fields line-of-sight loss (PL:28-65).  weights [N,S,1], termination_depth [N,1], steps [N,S,1]
    (sample mid-points in metres): inside +-sigma of the target depth the weights follow a Gaussian of std sigma / 3,
    in front of it they are pushed to zero; mean over rays with a valid, non-sky depth."""
    depth_mask = (termination_depth > 1.0) & (termination_depth < upper_bound)
    if sky_mask is not None:
        depth_mask = depth_mask & (sky_mask == 0.0)
    steps = steps.detach()
    td = termination_depth[:, None]
    std = sigma / URF_SIGMA_SCALE_FACTOR
    log_prob = -((steps - td) ** 2) / (2 * std ** 2) - math.log(std) - math.log(math.sqrt(2 * math.pi))
    near_mask = torch.logical_and(steps <= td + sigma, steps >= td - sigma)
    near = (near_mask * (weights - torch.exp(log_prob)) ** 2).sum(-2)
    empty = ((steps < td - sigma) * weights ** 2).sum(-2)
    return torch.mean((near + empty)[depth_mask])


def line_of_sight_sigma(step: int, start_step: int = 1000, end_step: int = 30000, max_sigma: float = 5.0,
                        min_sigma: float = 2.0) -> float:
    """MODEL:387-396."""
    frac = float(np.clip((step - start_step) / (end_step - start_step), 0.0, 1.0))
    return max_sigma - frac * (max_sigma - min_sigma)


def line_of_sight_mult(step: int, start_step: int = 1000, decay_steps: int = 5000, mult: float = 0.1) -> float:
    """MODEL:398-403."""
    if step <= start_step:
        return 0.0
    return mult / (2.0 ** (step // decay_steps))


def sky_blend(rgb_f: Tensor, acc_raw: Tensor, sem_f: Optional[Tensor], sky_rgb: Optional[Tensor],
              sky_sem: Optional[Tensor], training: bool = True):
    """Model epilogue (models/PreSight/nerfacto_nusc_ms.py:512-532): clamp the accumulation, add the sky colour /
    sky semantics behind the scene.  [N,3],[N,1],[N,C],[N,3],[N,C] -> (rgb, accumulation, semantics)."""
    accumulation = torch.clamp(acc_raw, min=0.0, max=1.0)
    rgb = rgb_f if training else torch.clamp(rgb_f, min=0.0, max=1.0)          # RN:221-229 (eval clamp)
    if sky_rgb is not None:
        rgb = rgb + (1.0 - accumulation) * sky_rgb
    sem = sem_f
    if sem_f is not None and sky_sem is not None:
        sem = sem_f + (1.0 - accumulation) * sky_sem
    return rgb, accumulation, sem


def rgb_loss(gt_rgb: Tensor, pred_rgb: Tensor) -> Tensor:
    """nn.MSELoss of models/PreSight/nerfacto_nusc_ms.py:314, 560-567."""
    return torch.mean((pred_rgb - gt_rgb) ** 2)


def sky_loss(accumulation: Tensor, sky_mask: Tensor, eps: float = 1e-7) -> Tensor:
    """model_components/PreSight/losses.py:106-115: BCE of the clipped accumulation against 1 - sky_mask
    (F.binary_cross_entropy clamps both logarithms at -100)."""
    target = 1.0 - sky_mask
    x = torch.clip(accumulation, min=eps, max=1 - eps)
    loss = -(target * torch.clamp(torch.log(x), min=-100.0) + (1.0 - target) * torch.clamp(torch.log(1.0 - x), min=-100.0))
    return loss.mean()


def semantic_loss(pred: Tensor, target: Tensor, clip: bool = True) -> Tensor:
    """model_components/PreSight/losses.py:117-125."""
    if clip:
        target = torch.clip(target, min=0.0, max=1.0)
    return torch.mean((pred - target) ** 2)


# --------------------------------------------------------------------------------------
# 8f-3  ray generation (the caller on the input side of the path)
# --------------------------------------------------------------------------------------
def pinhole_rays(c2w: Tensor, fx: Tensor, fy: Tensor, cx: Tensor, cy: Tensor, ray_indices: Tensor,
                 pixel_offset: float = 0.5) -> Dict[str, Tensor]:
    """Perspective ray generation without lens distortion: RayGenerator.forward (model_components/ray_generators.py:43-61)
    -> Cameras._generate_rays_from_coords (cameras/cameras.py:497-880, PERSPECTIVE branch :773-779).

    c2w [C,3,4], fx/fy/cx/cy [C], ray_indices [N,3] = (camera, row, col) -> origins [N,3], unit directions [N,3],
    pixel_area [N,1] (product of the distances to the directions of the +1-column and +1-row pixels), directions_norm."""
    cam, row, col = ray_indices[:, 0], ray_indices[:, 1], ray_indices[:, 2]
    y, x = row.float() + pixel_offset, col.float() + pixel_offset            # image_coords[y, x] (:311-312), (y, x) order
    fxr, fyr, cxr, cyr = fx[cam], fy[cam], cx[cam], cy[cam]
    coord = torch.stack([(x - cxr) / fxr, -(y - cyr) / fyr], -1)            # :613-615
    coord_x = torch.stack([(x - cxr + 1) / fxr, -(y - cyr) / fyr], -1)
    coord_y = torch.stack([(x - cxr) / fxr, -(y - cyr + 1) / fyr], -1)
    stack = torch.stack([coord, coord_x, coord_y], dim=0)                    # [3,N,2]
    dirs = torch.cat([stack, -torch.ones_like(stack[..., :1])], dim=-1)      # camera looks down -z (:777-779)
    rot = c2w[cam][:, :3, :3]
    dirs = torch.sum(dirs[..., None, :] * rot, dim=-1)                       # :844-846
    eps = torch.tensor([float(np.finfo(float).eps * 4.0)])                                              # camera_utils.py:30
    norm = torch.maximum(torch.linalg.vector_norm(dirs, dim=-1, keepdim=True), eps.to(dirs))            # camera_utils.py:299
    dirs = dirs / norm
    d = dirs[0]
    dx = torch.sqrt(torch.sum((d - dirs[1]) ** 2, dim=-1))                   # :854-855
    dy = torch.sqrt(torch.sum((d - dirs[2]) ** 2, dim=-1))
    return {"origins": c2w[cam][:, :3, 3], "directions": d, "pixel_area": (dx * dy)[..., None],
            "directions_norm": norm[0]}


# --------------------------------------------------------------------------------------
# a15-a18  compositing and renderers
# --------------------------------------------------------------------------------------
def get_weights(deltas: Tensor, densities: Tensor) -> Tensor:
    """[N,S,1],[N,S,1] -> weights [N,S,1].  Restates RAYS:138-150."""
    dd = deltas * densities
    alphas = 1 - torch.exp(-dd)
    acc = torch.cumsum(dd[..., :-1, :], dim=-2)
    acc = torch.cat([torch.zeros((*acc.shape[:1], 1, 1)), acc], dim=-2)
    return torch.nan_to_num(alphas * torch.exp(-acc))


def render_rgb(rgb: Tensor, weights: Tensor, background: Optional[Tensor] = None, training: bool = True) -> Tensor:
    """Σ w·rgb (+ bg·(1-Σw)).  Restates RN:102-117, 221-229."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    out = torch.sum(weights * rgb, dim=-2)
    if background is not None:
        out = out + background * (1.0 - torch.sum(weights, dim=-2))
    if not training:
        out = torch.clamp(out, 0.0, 1.0)
    return out


def render_accumulation(weights: Tensor) -> Tensor:
    """Restates RN:313."""
    return torch.sum(weights, dim=-2)


def render_depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """Σ w·t / (Σ w + 1e-10), clipped to the batch-global [min t, max t].  Restates RN:363-379."""
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
    return torch.clip(depth, steps.min(), steps.max())


def render_depth_threshold(weights: Tensor, starts: Tensor, ends: Tensor, threshold: float = 0.5) -> Tuple[Tensor, Tensor]:
    """First sample whose cumulative weight reaches `threshold`.  Restates RN:352-362.
    Returns (depth [N,1], index [N,1] int64)."""
    steps = (starts + ends) / 2
    cw = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * threshold
    idx = torch.searchsorted(cw, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=idx), idx


# --------------------------------------------------------------------------------------
# a13, a19-a21  sampler loop, model driver, prior query
# --------------------------------------------------------------------------------------
def _ms_prop_density(model: Model, level: int, pos: Tensor) -> Tensor:
    flat = pos.reshape(-1, 3)
    res = ms_dispatch(flat, model.centroids, lambda i, p: {"density": prop_density(model.props[level][i], p)},
                      len(model.props[level]))
    return res["density"].reshape(*pos.shape[:-1], 1)


def _ms_field_forward(model: Model, pos: Tensor, dirs: Tensor, app: Optional[Tensor]) -> Dict[str, Tensor]:
    shape = pos.shape[:-1]
    flat = pos.reshape(-1, 3)
    d = dirs.reshape(-1, 3)
    a = None if app is None else app.reshape(-1, app.shape[-1])
    res = ms_dispatch(flat, model.centroids, lambda i, p, dd, aa: ngp_forward(model.fields[i], p, dd, aa),
                      len(model.fields), extras=(d, a))
    return {k: v.reshape(*shape, -1) for k, v in res.items()}


def proposal_sample(model: Model, origins: Tensor, dirs: Tensor, nears: Tensor, fars: Tensor,
                    jitters: Optional[Sequence[Tensor]], anneal: float = 1.0, prop_grad: bool = True):
    """Level loop of the proposal sampler.  Restates RS:572-614.

    jitters: per level [N,1] uniform randoms (train, single jitter) or None for eval.
    Returns (final dict(sp_bins, eu_bins), weights_list, bins_list, inds_list).
    """
    cfg = model.cfg
    n = len(cfg.num_proposal_samples)
    fn, inv = spacing_fns(cfg.piecewise_thr)
    s_near, s_far = fn(nears), fn(fars)
    weights_list, bins_list, inds_list = [], [], []
    sp = eu = weights = None
    for lvl in range(n + 1):
        S = cfg.num_proposal_samples[lvl] if lvl < n else cfg.num_nerf_samples
        jit = None if jitters is None else jitters[lvl]
        if lvl == 0:
            sp, eu = spaced_bins(nears, fars, S, cfg.piecewise_thr, jit)
            inds_list.append(None)
        else:
            annealed = torch.pow(weights, anneal)
            sp, inds, _, _ = pdf_resample(annealed[..., 0], sp, S, jit, cfg.histogram_padding,
                                          eps=torch.finfo(torch.float32).eps)
            eu = inv(sp * s_far + (1 - sp) * s_near)
            inds_list.append(inds)
        if lvl < n:
            pos, starts, ends, deltas = sample_positions(origins, dirs, eu)
            if prop_grad:
                dens = _ms_prop_density(model, lvl, pos)
            else:
                with torch.no_grad():
                    dens = _ms_prop_density(model, lvl, pos)
            weights = get_weights(deltas, dens)
            weights_list.append(weights)
            bins_list.append((sp, eu))
    return (sp, eu), weights_list, bins_list, inds_list


def model_outputs(model: Model, origins: Tensor, dirs: Tensor, app: Optional[Tensor],
                  jitters: Optional[Sequence[Tensor]], anneal: float = 1.0, training: bool = True,
                  prop_grad: bool = True) -> Dict[str, object]:
    """One forward of the city NeRF.  Restates MODEL:452-546 (collider: scene_colliders.py:182-187).

    app: per-ray appearance embedding [N, A] (already looked up) or None.
    """
    cfg = model.cfg
    N = origins.shape[0]
    nears = torch.ones(N, 1) * (cfg.near if training else 0)
    fars = torch.ones(N, 1) * cfg.far
    (sp, eu), weights_list, bins_list, inds_list = proposal_sample(
        model, origins, dirs, nears, fars, jitters, anneal, prop_grad)
    pos, starts, ends, deltas = sample_positions(origins, dirs, eu)
    S = pos.shape[1]
    app_s = None if app is None else app[:, None, :].expand(N, S, app.shape[-1])
    fo = _ms_field_forward(model, pos, dirs[:, None, :].expand(N, S, 3), app_s)
    weights = get_weights(deltas, fo["density"])
    weights_list = weights_list + [weights]
    bins_list = bins_list + [(sp, eu)]
    rgb = render_rgb(fo["rgb"], weights, training=training)
    with torch.no_grad():
        depth, depth_idx = render_depth_threshold(weights, starts, ends)
    expected_depth = render_depth_expected(weights, starts, ends)
    acc = torch.clamp(render_accumulation(weights), 0.0, 1.0)
    out: Dict[str, object] = {}
    sky = None
    if model.sky is not None:
        res = ms_dispatch(origins, model.centroids, lambda i, p, dd, aa: sky_outputs(model.sky[i], dd, aa),
                          len(model.sky), extras=(dirs, app))
        sky = res
        rgb = rgb + (1.0 - acc) * sky["rgb"]
    out.update(rgb=rgb, accumulation=acc, depth=depth, depth_index=depth_idx, expected_depth=expected_depth)
    if "semantics" in fo:
        sem = torch.sum(fo["semantics"] * weights, dim=-2)
        if sky is not None and "semantics" in sky:
            sem = sem + (1.0 - acc) * sky["semantics"]
        out["semantics"] = sem
    out["weights_list"] = weights_list
    out["bins_list"] = bins_list
    out["inds_list"] = inds_list
    for i in range(len(cfg.num_proposal_samples)):
        eu_i = bins_list[i][1]
        out[f"prop_depth_{i}"] = render_depth_threshold(weights_list[i], eu_i[..., :-1, None], eu_i[..., 1:, None])[0]
    return out


def model_depth(model: Model, origins: Tensor, dirs: Tensor, threshold: float = 0.5) -> Dict[str, Tensor]:
    """Eval-mode depth for prior extraction.  Restates MODEL:688-708 (near=0 in eval)."""
    N = origins.shape[0]
    nears, fars = torch.zeros(N, 1), torch.ones(N, 1) * model.cfg.far
    with torch.no_grad():
        (sp, eu), _, _, _ = proposal_sample(model, origins, dirs, nears, fars, None, 1.0, prop_grad=False)
        pos, starts, ends, deltas = sample_positions(origins, dirs, eu)
        flat = pos.reshape(-1, 3)
        res = ms_dispatch(flat, model.centroids, lambda i, p: {"density": ngp_density(model.fields[i], p)[0]},
                          len(model.fields))
        w = get_weights(deltas, res["density"].reshape(N, -1, 1))
        d, _ = render_depth_threshold(w, starts, ends, threshold)
        return {"depth": d, "expected_depth": render_depth_expected(w, starts, ends)}


def prior_query(model: Model, points_scaled: Tensor) -> Tuple[Tensor, Tensor]:
    """Mean density over (proposal nets + main field) and clipped fp16 semantics at points.
    Restates XP:130-138."""
    with torch.no_grad():
        dens = []
        for lvl in range(len(model.props)):
            dens.append(_ms_prop_density(model, lvl, points_scaled).squeeze(-1))
        res = ms_dispatch(points_scaled, model.centroids,
                          lambda i, p: {"density": ngp_density(model.fields[i], p)[0]}, len(model.fields))
        dens.append(res["density"].squeeze(-1))
        mean = torch.stack(dens, dim=0).mean(dim=0)
        sem = ms_dispatch(points_scaled, model.centroids,
                          lambda i, p: {"semantics": ngp_semantics(model.fields[i], p)}, len(model.fields))
        feats = sem["semantics"].clip(0.0, 1.0).to(torch.float16)
    return mean, feats


# --------------------------------------------------------------------------------------
# bf16-operand emulation of the MLP (checker for the tensor-core path; not a reference restatement)
# --------------------------------------------------------------------------------------
def _bf(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class _Bf16Linear(torch.autograd.Function):
    """y = bf16(x) @ bf16(W)^T + b with fp32 accumulation; the backward rounds the incoming gradient and the
    saved operands to bf16 as well — the arithmetic of the bf16 MMA kernels (mlp_mma.cuh)."""

    @staticmethod
    def forward(ctx, x, w, b):
        xq, wq = _bf(x), _bf(w)
        ctx.save_for_backward(xq, wq)
        return xq @ wq.T + b

    @staticmethod
    def backward(ctx, g):
        xq, wq = ctx.saved_tensors
        gq = _bf(g)
        return gq @ wq, gq.T @ xq, gq.sum(dim=0)


def mlp_forward_bf16_emulated(x: Tensor, net: Mlp) -> Tensor:
    """Same network as mlp_forward (MLPF:157-174) with every matrix operand rounded to bf16."""
    n = len(net.weights)
    for i, (w, b) in enumerate(zip(net.weights, net.biases)):
        x = _Bf16Linear.apply(x, w, b)
        if i < n - 1:
            x = torch.relu(x)
    if net.out_act == "sigmoid":
        x = torch.sigmoid(x)
    return x
