"""Adam as PreSight configures it (reference: engine/optimizers.py:133-140 -> torch.optim.Adam with
configs/method_configs.py:115's lr 1e-2, eps 1e-15, weight_decay 1e-5), as ONE kernel per parameter (`ps_adam_step`).

The dense hash tables make the optimiser a bandwidth problem: torch's unfused update makes ~10 elementwise passes over
p / grad / m / v; for C2's 576 MiB of tables that is several GB of HBM traffic per step.  The fused step
reads p, grad, m, v once and writes p, m, v once.

Drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` (no amsgrad / maximize / capturable); learning-rate
schedulers work as usual (they edit `param_groups[i]["lr"]`)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import call, ptr, stream

GRAD_SCALE = 2.0 ** 10
"""The reference's trainer multiplies the loss by a GradScaler's initial scale (engine/trainer.py:70-73: use_grad_scaler
True, init_grad_scale 2^10, update_grad_scaler False) and then calls `optimizer.step()` WITHOUT unscaling
(trainer.py:481-486 -> optimizers.py:133-140): the optimiser sees gradients 1024 x larger.  Adam is invariant to that
except through eps (1e-15, negligible) and the coupled weight decay, which is added to the SCALED gradient and is thus
1024 x weaker than its nominal 1e-5.  To train like the reference, scale the loss the same way: `scale_loss(loss)`."""


def scale_loss(loss: torch.Tensor, scale: float = GRAD_SCALE) -> torch.Tensor:
    """loss * 2^10, what `GradScaler(init_scale=2**10).scale(loss)` returns (engine/trainer.py:481)."""
    return loss * scale


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


def shard_bounds(numel: int, rank: int, world: int, align: int = 4) -> Tuple[int, int, int]:
    """Contiguous shard of a flat parameter for `rank`: -> (lo, hi, padded shard length).  Every rank's shard has the same
    padded length (a multiple of `align` elements, so fp32 shards stay 16-byte aligned); the last shards may be shorter or
    empty in real elements."""
    per = (numel + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(rank * per, numel)
    return lo, min(lo + per, numel), per


class ShardedFusedAdam(torch.optim.Optimizer):
    """Adam over data-parallel replicas WITHOUT the dense all-reduce: per large parameter

        reduce-scatter(grad, AVG) -> ps_adam_step on this rank's 1/world shard -> all-gather(param)

    instead of DDP's all-reduce(grad) followed by a full Adam pass on every rank (my_pipeline.py:121-124 +
    optimizers.py:133-140).  Same bytes over NVLink (a ring all-reduce IS a reduce-scatter plus an all-gather), but each
    rank's optimiser pass touches 1/world of p / grad / m / v (7 x 576 MiB -> 7 x 72 MiB per step at 8 GPUs for C2) and
    keeps 1/world of the Adam state.  Parameters below `min_shard_numel` are averaged as one flat all-reduce and updated
    replicated.  The update arithmetic is `ps_adam_step`'s, i.e. torch.optim.Adam's; the result is the same on every rank.

    Call `step()` after backward INSTEAD of GradSynchronizer.finish() + optimizer.step().  Parameters without a gradient
    on a rank contribute zeros (DDP's find_unused_parameters=True behaviour)."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, group: Optional["dist.ProcessGroup"] = None, min_shard_numel: int = 1 << 16) -> None:
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.min_shard_numel = min_shard_numel
        self._nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def _adam(self, p_flat, g_flat, state, group) -> None:
        beta1, beta2 = group["betas"]
        with torch.cuda.device(p_flat.device):
            call("ps_adam_step", ptr(p_flat), ptr(g_flat), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), p_flat.numel(),
                 float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                 int(state["step"]), stream())

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None, "ShardedFusedAdam does not take a closure"
        W = self.world
        for group in self.param_groups:
            big = [p for p in group["params"] if p.numel() >= self.min_shard_numel and W > 1]
            small = [p for p in group["params"] if not (p.numel() >= self.min_shard_numel and W > 1)]
            pending = []
            for p in big:
                assert p.dtype == torch.float32 and p.is_contiguous(), "ShardedFusedAdam needs contiguous fp32 parameters"
                lo, hi, per = shard_bounds(p.numel(), self.rank, W)
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros(per, device=p.device, dtype=torch.float32)
                    state["exp_avg_sq"] = torch.zeros(per, device=p.device, dtype=torch.float32)
                    state["shard"] = torch.zeros(per, device=p.device, dtype=torch.float32)
                state["step"] += 1
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                gflat = g.reshape(-1)
                if per * W != p.numel():                         # pad to W equal shards
                    gpad = torch.zeros(per * W, device=p.device, dtype=torch.float32)
                    gpad[:p.numel()].copy_(gflat)
                    gflat = gpad
                gshard = torch.empty(per, device=p.device, dtype=torch.float32)
                if self._nccl:
                    work = dist.reduce_scatter_tensor(gshard, gflat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
                else:                                            # backends without reduce-scatter (gloo, CPU tests)
                    dist.all_reduce(gflat, group=self.group)
                    gshard.copy_(gflat[self.rank * per:(self.rank + 1) * per] / W)
                    work = None
                pending.append((p, state, gshard, lo, hi, per, work))
            # small parameters: one flat all-reduce, replicated update (overlaps the reduce-scatters above)
            live = [p for p in small if p.grad is not None or W > 1]
            if live and W > 1:
                flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in live])
                dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM, group=self.group)
                if not self._nccl:
                    flat.div_(W)
                off = 0
                for p in live:
                    n = p.numel()
                    if p.grad is None:
                        p.grad = torch.empty_like(p)
                    p.grad.copy_(flat[off:off + n].view_as(p))
                    off += n
            for p in live:
                if p.grad is None:
                    continue
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                self._adam(p.data.view(-1), p.grad.contiguous().view(-1), state, group)
            # shards: Adam on the owned slice, then all-gather the updated parameter
            gathers = []
            for p, state, gshard, lo, hi, per, work in pending:
                if work is not None:
                    work.wait()
                shard = state["shard"]
                pflat = p.data.view(-1)
                if hi > lo:
                    shard[:hi - lo].copy_(pflat[lo:hi])
                self._adam(shard, gshard, state, group)
                if per * W == p.numel():
                    out = pflat
                else:
                    out = torch.empty(per * W, device=p.device, dtype=torch.float32)
                if self._nccl:
                    w2 = dist.all_gather_into_tensor(out, shard, group=self.group, async_op=True)
                else:
                    chunks = [torch.empty_like(shard) for _ in range(W)]
                    dist.all_gather(chunks, shard, group=self.group)
                    out.copy_(torch.cat(chunks))
                    w2 = None
                gathers.append((p, out, w2))
            for p, out, w2 in gathers:
                if w2 is not None:
                    w2.wait()
                if out.data_ptr() != p.data.data_ptr():
                    p.data.view(-1).copy_(out[:p.numel()])
        return None
