#!/usr/bin/env python
"""Golden fixture for the distortion loss from the LIVE reference (build container only).

    python tests/golden/make_golden_distortion.py      # rewrites tests/golden/distortion.npz

Runs the reference's own `distortion_loss` / `lossfun_distortion` (model_components/losses.py:130-149) on seeded inputs
shaped like the final level of the three sample-count regimes (64, 48 and a ragged 5), with tied bin edges and exact-zero
weights, and stores the inputs, the per-ray terms, the scalar loss and its gradient w.r.t. the weights.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shims and sys.path for the reference)
import torch  # noqa: E402
from make_golden_losses import bins, weights  # noqa: E402
from nerfstudio.model_components import losses as RL  # noqa: E402


def main():
    g = torch.Generator().manual_seed(777)
    out = {}
    for name, (n, s) in {"a": (96, 64), "b": (33, 48), "c": (7, 5), "d": (5, 200)}.items():
        c = bins(g, n, s)
        if name == "a":
            c[:8, 1:-1] = bins(g, 8, s)[:, 1:-1].round(decimals=2).sort(dim=-1).values   # tied edges
        w = weights(g, n, s)[..., None].requires_grad_(True)                              # [N,S,1] like weights_list[-1]

        class RS:      # the two attributes ray_samples_to_sdist reads (losses.py:100-105)
            def __init__(self, b):
                self.spacing_starts, self.spacing_ends = b[:, :-1, None], b[:, 1:, None]

        loss = RL.distortion_loss([w], [RS(c)])
        loss.backward()
        out[f"{name}/c"], out[f"{name}/w"], out[f"{name}/loss"], out[f"{name}/g"] = c, w.detach(), loss, w.grad
        out[f"{name}/terms"] = RL.lossfun_distortion(c, w.detach()[..., 0])
    MG.save("distortion.npz", out)


if __name__ == "__main__":
    main()
