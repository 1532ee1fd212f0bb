"""Adam as PreSight configures it (reference: engine/optimizers.py:133-140 -> torch.optim.Adam with
configs/method_configs.py:115's lr 1e-2, eps 1e-15, weight_decay 1e-5), as ONE kernel per parameter (`ps_adam_step`).

The dense hash tables make the optimiser a bandwidth problem: torch's unfused update makes ~10 elementwise passes over
p / grad / m / v; for C2's 576 MiB of tables that is several GB of HBM traffic per step.  The fused step
reads p, grad, m, v once and writes p, m, v once.

Drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` (no amsgrad / maximize / capturable); learning-rate
schedulers work as usual (they edit `param_groups[i]["lr"]`)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch

from ._lib import call, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0) -> None:
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")           # the checks of torch.optim.Adam.__init__
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 parameters")
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if grad.dtype != torch.float32:
                    grad = grad.float()
                with torch.cuda.device(p.device):            # the launch goes to the parameter's device and its current stream
                    call("ps_adam_step", ptr(p), ptr(grad), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), p.numel(),
                         float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                         int(state["step"]), stream())
        return loss
