"""Samplers (reference: nerfstudio/model_components/ray_samplers.py) on the b200 sampler kernels.

`SpacedSampler` here is PreSight's piecewise sampler (nerfacto_nusc_ms.py:312-317): the spacing function is
given by its threshold instead of two Python callables so that it can run inside the kernel.
Random jitter is drawn with `torch.rand` on the ray tensors' device, exactly where the reference draws it
(ray_samplers.py:105, 322), and handed to the kernels as an input.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..cameras.rays import RayBundle, RaySamples


def piecewise_spacing(thr: float):
    """The reference's spacing_fn / spacing_fn_inv pair (nerfacto_nusc_ms.py:312-317), kept for API parity."""
    def fn(x):
        return torch.where(x < thr, x / (2 * thr), 1 - 1 / (2 * x / thr))

    def inv(x):
        return torch.where(x < 0.5, x * (2 * thr), thr / (2 - 2 * x))
    return fn, inv


class PiecewiseSpacingToEuclidean:
    """Callable stored in `RaySamples.spacing_to_euclidean_fn` (ray_samplers.py:113-114).  Calling it evaluates the
    mapping with torch ops like the reference's closure; the kernels read `.thr` and evaluate it on chip."""

    def __init__(self, thr: float, nears: Tensor, fars: Tensor) -> None:
        self.thr = float(thr)
        self.nears, self.fars = nears, fars

    def __call__(self, x: Tensor) -> Tensor:
        fn, inv = piecewise_spacing(self.thr)
        s_near, s_far = fn(self.nears), fn(self.fars)
        return inv(x * s_far + (1 - x) * s_near)


class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)


class SpacedSampler(Sampler):
    """Initial sampler (ray_samplers.py:49-128) with PreSight's piecewise spacing of threshold `thr`."""

    def __init__(self, piecewise_threshold: float, num_samples: Optional[int] = None, train_stratified: bool = True,
                 single_jitter: bool = True) -> None:
        super().__init__(num_samples=num_samples)
        if not single_jitter:
            raise NotImplementedError("PreSight uses single_jitter=True; per-sample jitter is not implemented")
        self.thr = float(piecewise_threshold)
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn, self.spacing_fn_inv = piecewise_spacing(self.thr)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None,
                             t_rand: Optional[Tensor] = None) -> RaySamples:
        assert ray_bundle is not None
        assert ray_bundle.nears is not None
        assert ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        num_rays = ray_bundle.origins.shape[0]
        if self.train_stratified and self.training:
            if t_rand is None:
                t_rand = torch.rand((num_rays, 1), dtype=torch.float32, device=ray_bundle.origins.device)
        else:
            t_rand = None
        sp, eu = ops.spaced_bins(ray_bundle.nears, ray_bundle.fars, num_samples, self.thr, t_rand)
        return RayBundle.samples_from_bins(ray_bundle, sp, eu,
                                           PiecewiseSpacingToEuclidean(self.thr, ray_bundle.nears, ray_bundle.fars))


class PDFSampler(Sampler):
    """Inverse-CDF sampler (ray_samplers.py:244-372) on `ps_pdf_resample`."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        if include_original:
            raise NotImplementedError("include_original=True (sort path) is not used by ProposalNetworkSampler")
        if not single_jitter:
            raise NotImplementedError("PreSight uses single_jitter=True; per-sample jitter is not implemented")
        self.train_stratified = train_stratified
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.single_jitter = single_jitter

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None,
                             weights: Optional[Tensor] = None, num_samples: Optional[int] = None, eps: float = 1e-5,
                             rand: Optional[Tensor] = None, anneal: float = 1.0) -> RaySamples:
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None, "weights must be provided"
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        assert (ray_samples.spacing_starts is not None and ray_samples.spacing_ends is not None
                ), "ray_sample spacing_starts and spacing_ends must be provided"
        assert ray_samples.spacing_to_euclidean_fn is not None, "ray_samples.spacing_to_euclidean_fn must be provided"
        to_eu = ray_samples.spacing_to_euclidean_fn
        if not isinstance(to_eu, PiecewiseSpacingToEuclidean):
            raise NotImplementedError("the b200 PDF sampler needs samples produced by the piecewise SpacedSampler")
        thr = to_eu.thr
        existing = ray_samples.sp_bins
        if existing is None:
            existing = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)
        N = weights.shape[0]
        if self.train_stratified and self.training:
            if rand is None:
                rand = torch.rand((N, 1), device=weights.device)
        else:
            rand = None
        sp, eu = ops.pdf_resample(weights[..., 0], existing, num_samples, rand, ray_bundle.nears, ray_bundle.fars, thr,
                                  padding=self.histogram_padding, eps=eps, anneal=anneal)
        return RayBundle.samples_from_bins(ray_bundle, sp, eu, to_eu)


class ProposalNetworkSampler(Sampler):
    """Proposal sampling loop (ray_samplers.py:523-614)."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = True,
                 update_sched: Callable = lambda x: 1, initial_sampler: Optional[Sampler] = None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        if initial_sampler is None:
            raise NotImplementedError("pass the piecewise SpacedSampler as initial_sampler (PreSight's configuration)")
        self.initial_sampler = initial_sampler
        self.pdf_sampler = PDFSampler(include_original=False, single_jitter=single_jitter)
        self.use_fused = True      # take the level-fused fast path when a density fn's owner offers it
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, density_fns: Optional[List[Callable]] = None,
                             jitters: Optional[List[Tensor]] = None) -> Tuple[RaySamples, List, List]:
        assert ray_bundle is not None
        assert density_fns is not None
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights = None
        ray_samples = None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            jit = None if jitters is None else jitters[i_level]
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples, t_rand=jit)
            else:
                assert weights is not None
                # the annealing pow of ray_samplers.py:597 is applied inside the kernel
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples,
                                               eps=torch.finfo(torch.float32).eps, rand=jit, anneal=self._anneal)
            if is_prop:
                owner = getattr(density_fns[i_level], "__self__", None)
                fused_ok = self.use_fused and owner is not None and getattr(owner, "supports_fused", lambda: False)()
                if fused_ok:
                    # level-fused fast path: same maths, one autograd node (presight_b200/fused.py)
                    with torch.set_grad_enabled(updated and torch.is_grad_enabled()):
                        weights = owner.level_weights(ray_bundle.origins, ray_bundle.directions,
                                                      ray_samples.frustums.eu_bins)
                else:
                    if updated:
                        density = density_fns[i_level](ray_samples.frustums.get_positions())
                    else:
                        with torch.no_grad():
                            density = density_fns[i_level](ray_samples.frustums.get_positions())
                    weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list
