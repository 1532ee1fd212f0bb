#!/usr/bin/env python
"""Golden fixture for the per-ray tail of the train step from the LIVE reference (build container only).

    python tests/golden/make_golden_render_losses.py      # rewrites tests/golden/render_losses.npz

The loss terms are the reference's own functions — `sky_loss` / `semantic_loss`
(model_components/PreSight/losses.py:106-125) and `MSELoss` (models/PreSight/nerfacto_nusc_ms.py:314, 560-567); the
model epilogue (nerfacto_nusc_ms.py:512-532) is restated here line by line because the model module does not import in
this container.  Inputs are seeded and include the edge cases: accumulation below 0 / above 1 / exactly 0
and 1 / within eps of both ends, sky masks of 0 and 1, target features outside [0, 1].
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
    g = torch.Generator().manual_seed(4321)
    out = {}
    for name, (n, c) in {"a": (257, 64), "b": (31, 8)}.items():
        rgb_f = torch.rand(n, 3, generator=g) * 1.2 - 0.1
        acc_raw = torch.rand(n, 1, generator=g) * 1.3 - 0.15
        acc_raw[:8, 0] = torch.tensor([0.0, 1.0, -0.25, 1.5, 5e-8, 1 - 5e-8, 1e-7, 0.5])
        sem_f = torch.randn(n, c, generator=g) * 0.5 + 0.3
        sky_rgb = torch.rand(n, 3, generator=g)
        sky_sem = torch.randn(n, c, generator=g) * 0.4
        gt_rgb = torch.rand(n, 3, generator=g)
        sky = (torch.rand(n, 1, generator=g) < 0.3).float()
        gt_sem = torch.randn(n, c, generator=g) * 0.6 + 0.4          # partly outside [0, 1]
        leaves = [t.requires_grad_(True) for t in (rgb_f, acc_raw, sem_f, sky_rgb, sky_sem)]
        # ---- epilogue, nerfacto_nusc_ms.py:512-532 (training mode)
        accumulation = torch.clamp(acc_raw, min=0.0, max=1.0)
        rgb = rgb_f + (1.0 - accumulation) * sky_rgb
        semantics = sem_f + (1.0 - accumulation) * sky_sem
        # ---- loss terms, nerfacto_nusc_ms.py:560-576, 641-645
        l_rgb = torch.nn.MSELoss()(gt_rgb, rgb)
        l_sky = PL.sky_loss(accumulation.view(-1, 1), sky.view(-1, 1))
        l_sem = PL.semantic_loss(semantics, gt_sem)
        total = l_rgb + 0.001 * l_sky + 0.5 * l_sem
        total.backward()
        out.update({f"{name}/rgb_f": rgb_f, f"{name}/acc_raw": acc_raw, f"{name}/sem_f": sem_f,
                    f"{name}/sky_rgb": sky_rgb, f"{name}/sky_sem": sky_sem, f"{name}/gt_rgb": gt_rgb,
                    f"{name}/sky": sky, f"{name}/gt_sem": gt_sem, f"{name}/rgb": rgb, f"{name}/acc": accumulation,
                    f"{name}/sem": semantics, f"{name}/losses": torch.stack([l_rgb, l_sky, l_sem]),
                    f"{name}/total": total})
        for k, t in zip(("rgb_f", "acc_raw", "sem_f", "sky_rgb", "sky_sem"), leaves):
            out[f"{name}/g_{k}"] = t.grad
    MG.save("render_losses.npz", out)


if __name__ == "__main__":
    main()
