"""Host-side learning-rate schedules used by the ASR recipes (scalar arithmetic, not a kernel).
  noam       espresso/optim/lr_scheduler/noam_lr_scheduler.py:40-78
  tri_stage  fairseq/optim/lr_scheduler/tri_stage_lr_scheduler.py:50-175
"""
import math

from ..registry import register_lr_scheduler


@register_lr_scheduler("noam")
class NoamLRScheduler:
    """lr = factor * model_size^-0.5 * min(n^-0.5, n * warmup^-1.5); after warm-up never below final_lr.
    `lr` in the recipe YAML (e.g. 5.0) is the factor (noam_lr_scheduler.py:52-63)."""

    def __init__(self, lr, warmup_steps, model_size, final_lr=None):
        self.factor = lr
        self.warmup_steps = max(1, warmup_steps)
        self.model_size = model_size
        self.final_lr = final_lr
        self.lr = self.step_update(0)

    def step_update(self, num_updates):
        n = num_updates + 1
        lr = self.factor * self.model_size ** -0.5 * min(n ** -0.5, n * self.warmup_steps ** -1.5)
        if self.final_lr is not None and n > self.warmup_steps:
            lr = max(lr, self.final_lr)
        self.lr = lr
        return lr


@register_lr_scheduler("tri_stage")
class TriStageLRScheduler:
    def __init__(self, lr, warmup_steps, hold_steps, decay_steps, init_lr_scale=0.01, final_lr_scale=0.01):
        self.peak_lr = lr
        self.init_lr = init_lr_scale * lr
        self.final_lr = final_lr_scale * lr
        self.warmup_steps, self.hold_steps, self.decay_steps = warmup_steps, hold_steps, decay_steps
        self.warmup_rate = (self.peak_lr - self.init_lr) / warmup_steps if warmup_steps != 0 else 0
        self.decay_factor = -math.log(final_lr_scale) / decay_steps
        self.lr = self.init_lr

    def step_update(self, num_updates):
        n = num_updates
        if n < self.warmup_steps:
            lr = self.init_lr + self.warmup_rate * n
        elif n < self.warmup_steps + self.hold_steps:
            lr = self.peak_lr
        elif n <= self.warmup_steps + self.hold_steps + self.decay_steps:
            lr = self.peak_lr * math.exp(-self.decay_factor * (n - self.warmup_steps - self.hold_steps))
        else:
            lr = self.final_lr
        self.lr = lr
        return lr
