"""Learning-rate schedule PreSight trains with (reference: nerfstudio/engine/my_schedulers.py:34-70,
configs/method_configs.py:116-119): linear warm-up from 1 % over `warmup_steps`, then x0.33 at every milestone.

The reference builds it from torch's own schedulers (ChainedScheduler[LinearLR, MultiStepLR]) and so does this module — it
is host-side logic, one scalar per step; `lr_factor` is the closed form, used by the tests and by callers that drive
`FusedAdam` / `ShardedFusedAdam` param groups by hand.  Note the reference's quirk, kept: `gamma` of the config is ignored,
the decay factor is hard-coded to 0.33 (my_schedulers.py:65)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

from torch.optim import Optimizer, lr_scheduler


@dataclass
class WarmupMultiStepSchedulerConfig:
    """my_schedulers.py:34-47."""
    max_steps: int = 1000000
    gamma: float = 0.33            # unused by the reference as well (see module docstring)
    milestones: Tuple[int, ...] = (500000, 750000, 900000)
    warmup_steps: Optional[int] = None

    def setup(self) -> "WarmupMultiStepScheduler":
        return WarmupMultiStepScheduler(self)


class WarmupMultiStepScheduler:
    """my_schedulers.py:50-70."""

    def __init__(self, config: WarmupMultiStepSchedulerConfig) -> None:
        self.config = config

    def get_scheduler(self, optimizer: Optimizer, lr_init: float):
        return lr_scheduler.ChainedScheduler([
            lr_scheduler.LinearLR(optimizer, start_factor=0.01, total_iters=self.config.warmup_steps),
            lr_scheduler.MultiStepLR(optimizer, milestones=list(self.config.milestones), gamma=0.33),
        ])

    def lr_factor(self, step: int) -> float:
        """lr(step) / lr_init after `step` calls of scheduler.step()."""
        w = self.config.warmup_steps
        warm = 1.0 if (w is None or w <= 0 or step >= w) else 0.01 + (1.0 - 0.01) * step / w
        return warm * 0.33 ** sum(1 for m in self.config.milestones if step >= m)


def presight_scheduler(max_iterations: int = 100000) -> WarmupMultiStepSchedulerConfig:
    """The schedule of every shipped method config (method_configs.py:116-119): milestones at 1/4, 1/2, 3/4 of the run,
    warm-up over the first tenth."""
    return WarmupMultiStepSchedulerConfig(max_steps=max_iterations,
                                          milestones=(max_iterations // 4, max_iterations // 2, max_iterations * 3 // 4),
                                          warmup_steps=max_iterations // 10)
