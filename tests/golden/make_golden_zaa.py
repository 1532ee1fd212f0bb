#!/usr/bin/env python
"""Golden fixture for the z-anti-aliased (zip-NeRF) interlevel loss — the reference's default proposal loss
(`enable_z_anti_aliasing=True`, models/PreSight/nerfacto_nusc_ms.py:129,293-295) — from the LIVE reference.

    python tests/golden/make_golden_zaa.py      # rewrites tests/golden/zaa.npz

Runs `z_anti_anliasing_interlevel_loss`, `blur_stepfun` and `sorted_interp_quad`
(model_components/PreSight/losses.py:127-206) with the config's pulse widths (0.03, 0.003) on seeded inputs shaped like
C2 (128/64 proposal samples, 64 final), C1 (256/96, 48) and a ragged small case; stores the inputs, the blurred
histograms, the interpolated envelopes w_s, the scalar loss and its gradient w.r.t. the proposal weights.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shims and sys.path for the reference)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from make_golden_losses import bins, weights  # noqa: E402
from nerfstudio.model_components.PreSight import losses as PL  # noqa: E402

PULSE = (0.03, 0.003)


def main():
    g = torch.Generator().manual_seed(2024)
    out = {"pulse_width": np.asarray(PULSE, dtype=np.float64)}
    for name, (n, s, sps) in {"a": (64, 64, (128, 64)), "b": (21, 48, (256, 96)), "c": (7, 5, (3, 9))}.items():
        c, w = bins(g, n, s), weights(g, n, s)
        if name == "a":
            c[:6, 1:-1] = bins(g, 6, s)[:, 1:-1].round(decimals=2).sort(dim=-1).values     # tied edges (zero-width bins)
            c[:6] = c[:6] + torch.arange(s + 1) * 1e-6                                      # ... kept strictly increasing
            c[:, 0], c[:, -1] = 0.0, 1.0
        ws = [weights(g, n, sp).requires_grad_(True) for sp in sps]
        ts = [bins(g, n, sp) for sp in sps]

        class RS:      # the two attributes ray_samples_to_sdist reads (model_components/losses.py:100-105)
            def __init__(self, b):
                self.spacing_starts, self.spacing_ends = b[:, :-1, None], b[:, 1:, None]

        loss = PL.z_anti_anliasing_interlevel_loss([x[..., None] for x in ws] + [w[..., None]],
                                                  [RS(t) for t in ts] + [RS(c)], PULSE)
        loss.backward()
        out[f"{name}/c"], out[f"{name}/w"], out[f"{name}/loss"] = c, w, loss
        wn = w / (c[..., 1:] - c[..., :-1])
        for i, (t, x) in enumerate(zip(ts, ws)):
            ci, wi = PL.blur_stepfun(c, wn, PULSE[i])
            area = 0.5 * (wi[..., 1:] + wi[..., :-1]) * (ci[..., 1:] - ci[..., :-1])
            cdfs = torch.cat([torch.zeros_like(area[..., :1]), torch.cumsum(area, dim=-1)], dim=-1)
            w_s = torch.diff(PL.sorted_interp_quad(t, ci, wi, cdfs), dim=-1)
            out[f"{name}/t{i}"], out[f"{name}/w{i}"], out[f"{name}/g{i}"] = t, x.detach(), x.grad
            out[f"{name}/xr{i}"], out[f"{name}/yr{i}"], out[f"{name}/ws{i}"] = ci, wi, w_s
        out[f"{name}/n_levels"] = np.int64(len(sps))
    MG.save("zaa.npz", out)


if __name__ == "__main__":
    main()
