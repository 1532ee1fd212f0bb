#!/usr/bin/env python
"""Golden fixture for the loss stack from the LIVE reference (build container only).

    python tests/golden/make_golden_losses.py      # rewrites tests/golden/losses.npz

Runs the reference's own `lossfun_outer` / `interlevel_loss` arithmetic (model_components/losses.py:47-126) on
seeded, realistically shaped inputs (sorted spacing-domain bins, normalised weights with exact zeros and ties) and
stores inputs, the per-sample loss terms, the scalar loss and its gradient w.r.t. the proposal weights.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shims and sys.path for the reference)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from nerfstudio.model_components import losses as RL  # noqa: E402


def bins(g, n, s):
    edges = torch.rand(n, s + 1, generator=g).sort(dim=-1).values
    edges[:, 0], edges[:, -1] = 0.0, 1.0
    return edges


def weights(g, n, s):
    w = torch.rand(n, s, generator=g) ** 4
    w[torch.rand(n, s, generator=g) < 0.2] = 0.0          # exact zeros
    return w / (w.sum(-1, keepdim=True) + 1e-3)


def main():
    g = torch.Generator().manual_seed(1234)
    out = {}
    for name, (n, s, sps) in {"a": (96, 64, (128, 64)), "b": (33, 48, (256, 96)), "c": (7, 5, (3,))}.items():
        c, w = bins(g, n, s), weights(g, n, s)
        if name == "a":
            c[:8, 1:-1] = bins(g, 8, s)[:, 1:-1].round(decimals=2).sort(dim=-1).values   # tied edges
        ws = [weights(g, n, sp).requires_grad_(True) for sp in sps]
        ts = [bins(g, n, sp) for sp in sps]
        if name == "a":
            ts[1][:, ::2] = c[:, ::2][:, : ts[1][:, ::2].shape[1]]                         # coincident edges
            ts[1] = ts[1].sort(dim=-1).values

        class RS:      # the two attributes ray_samples_to_sdist reads (losses.py:100-105)
            def __init__(self, b):
                self.spacing_starts, self.spacing_ends = b[:, :-1, None], b[:, 1:, None]

        loss = RL.interlevel_loss([x[..., None] for x in ws] + [w[..., None]], [RS(t) for t in ts] + [RS(c)])
        loss.backward()
        out[f"{name}/c"], out[f"{name}/w"], out[f"{name}/loss"] = c, w, loss
        for i, (t, x) in enumerate(zip(ts, ws)):
            out[f"{name}/t{i}"], out[f"{name}/w{i}"], out[f"{name}/g{i}"] = t, x.detach(), x.grad
            out[f"{name}/terms{i}"] = RL.lossfun_outer(c, w, t, x.detach())
        out[f"{name}/n_levels"] = np.int64(len(sps))
    MG.save("losses.npz", out)


if __name__ == "__main__":
    main()
