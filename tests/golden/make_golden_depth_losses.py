#!/usr/bin/env python
"""Golden fixture for the depth-supervision terms of the loss dict from the LIVE reference.

    python tests/golden/make_golden_depth_losses.py      # rewrites tests/golden/depth_losses.npz

Runs the reference's own `expected_monodepth_loss` (normalised and inverse), `expected_depth_loss` and
`line_of_sight_loss` (model_components/PreSight/losses.py:28-103) on seeded inputs: depths on both sides of the validity
window (<= 1 m, >= upper bound), sky rays, sample mid-points straddling the +-sigma band; stores inputs, losses and the
gradients w.r.t. the predicted depth / the weights.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shims and sys.path for the reference)
import torch  # noqa: E402
from nerfstudio.model_components.PreSight import losses as PL  # noqa: E402


def main():
    g = torch.Generator().manual_seed(99)
    out = {}
    n, s = 193, 64
    depth = torch.rand(n, 1, generator=g) * 60.0
    depth[:6, 0] = torch.tensor([0.5, 1.0, 39.999, 40.0, 55.0, 20.0])
    sky = (torch.rand(n, 1, generator=g) < 0.25).float()
    pred = (depth + torch.randn(n, 1, generator=g) * 4.0).clamp_min(0.01).requires_grad_(True)
    steps = torch.sort(torch.rand(n, s, 1, generator=g) * 70.0, dim=1).values
    w = (torch.rand(n, s, 1, generator=g) ** 3 * 0.2).requires_grad_(True)
    out.update(depth=depth, sky=sky, pred=pred.detach(), steps=steps, w=w.detach())
    for name, kw in {"mono": dict(upper_bound=40.0, inverse=False), "mono_inv": dict(upper_bound=40.0, inverse=True)}.items():
        pred.grad = None
        loss = PL.expected_monodepth_loss(termination_depth=depth, predicted_depth=pred, sky_mask=sky, **kw)
        loss.backward()
        out[f"{name}/loss"], out[f"{name}/g"] = loss, pred.grad.clone()
    pred.grad = None
    loss = PL.expected_depth_loss(termination_depth=depth, predicted_depth=pred, upper_bound=75.0)
    loss.backward()
    out["lidar/loss"], out["lidar/g"] = loss, pred.grad.clone()
    for name, (sigma, use_sky, ub) in {"los_a": (5.0, True, 40.0), "los_b": (2.0, False, 75.0)}.items():
        w.grad = None
        loss = PL.line_of_sight_loss(weights=w, termination_depth=depth, steps=steps, sigma=sigma,
                                     sky_mask=sky if use_sky else None, upper_bound=ub)
        loss.backward()
        out[f"{name}/loss"], out[f"{name}/g"] = loss, w.grad.clone()
    MG.save("depth_losses.npz", out)


if __name__ == "__main__":
    main()
