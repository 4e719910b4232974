"""Host-side learning-rate schedules used by the ASR recipes (scalar arithmetic, not a kernel).
  noam       espresso/optim/lr_scheduler/noam_lr_scheduler.py:40-78
  tri_stage  fairseq/optim/lr_scheduler/tri_stage_lr_scheduler.py:50-175
  reduce_lr_on_plateau_v2  espresso/optim/lr_scheduler/reduce_lr_on_plateau_v2.py:30-60 (+ fairseq's reduce_lr_on_plateau, torch's plateau rule)
  polynomial_decay_v2      espresso/optim/lr_scheduler/polynomial_decay_schedule.py:16-20 (+ fairseq's polynomial_decay)
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


@register_lr_scheduler("reduce_lr_on_plateau_v2")
class ReduceLROnPlateauV2LRScheduler:
    """The `speech_lstm` recipes' schedule (espresso/optim/lr_scheduler/reduce_lr_on_plateau_v2.py:30-60 over
    fairseq/optim/lr_scheduler/reduce_lr_on_plateau.py:55-143): optional linear warm-up per update; at the end of every epoch
    the learning rate is multiplied by `lr_shrink` when the validation metric has not improved by more than the relative
    `lr_threshold` for more than `lr_patience` epochs (the plateau rule of torch.optim.lr_scheduler.ReduceLROnPlateau, mode
    'rel', no cool-down), never below final_lr_scale * lr; before `start_reduce_lr_epoch` the rate stays at lr."""

    def __init__(self, lr, lr_shrink=0.1, lr_threshold=1e-4, lr_patience=0, warmup_updates=0, warmup_init_lr=-1.0,
                 start_reduce_lr_epoch=0, final_lr_scale=0.01, maximize_best_checkpoint_metric=False):
        self.base_lr = lr
        self.factor, self.threshold, self.patience = lr_shrink, lr_threshold, lr_patience
        self.start_epoch = start_reduce_lr_epoch
        self.min_lr = final_lr_scale * lr
        self.maximize = maximize_best_checkpoint_metric
        self.best = -math.inf if self.maximize else math.inf
        self.num_bad_epochs = 0
        self.last_epoch = 0
        self.warmup_updates = warmup_updates
        if warmup_init_lr < 0:
            warmup_init_lr = 0.0 if warmup_updates > 0 else lr
        self.warmup_init_lr = warmup_init_lr
        self.lr_step = (lr - warmup_init_lr) / warmup_updates if warmup_updates > 0 else 0.0
        self.warmup_end = warmup_updates <= 0
        self.lr = lr if self.warmup_end else warmup_init_lr

    def _is_better(self, a):
        if self.maximize:
            return a > self.best * (self.threshold + 1.0)
        return a < self.best * (1.0 - self.threshold)

    def step(self, epoch, val_loss=None):
        """End of `epoch` (1-based, as fairseq counts them) with the validation metric of that epoch."""
        if epoch < self.start_epoch:
            self.last_epoch = epoch
            self.lr = self.base_lr
            return self.lr
        if val_loss is not None and self.warmup_end:
            self.last_epoch += 1
            if self._is_better(float(val_loss)):
                self.best = float(val_loss)
                self.num_bad_epochs = 0
            else:
                self.num_bad_epochs += 1
            if self.num_bad_epochs > self.patience:
                new_lr = max(self.lr * self.factor, self.min_lr)
                if self.lr - new_lr > 1e-8:
                    self.lr = new_lr
                self.num_bad_epochs = 0
        else:
            self.last_epoch = epoch
        return self.lr

    def step_update(self, num_updates):
        if self.warmup_updates > 0:
            if num_updates <= self.warmup_updates:
                self.lr = self.warmup_init_lr + num_updates * self.lr_step
            elif not self.warmup_end:
                self.warmup_end = True
        return self.lr

    def state_dict(self):
        return {"best": self.best, "last_epoch": self.last_epoch}

    def load_state_dict(self, state):
        self.best = state["best"]
        if "last_epoch" in state:
            self.last_epoch = state["last_epoch"]


@register_lr_scheduler("polynomial_decay_v2")
class PolynomialDecayV2LRScheduler:
    """espresso/optim/lr_scheduler/polynomial_decay_schedule.py:16-20 over fairseq's polynomial_decay
    (fairseq/optim/lr_scheduler/polynomial_decay_schedule.py:40-96): linear warm-up to lr, then
    (lr - end) * (1 - (n - warmup) / (total - warmup))^power + end, `end` from total_num_update on; the per-epoch hook of the
    base class is a no-op in the v2 variant."""

    def __init__(self, lr, total_num_update, warmup_updates=0, end_learning_rate=0.0, power=1.0):
        assert total_num_update > 0
        self.base_lr, self.total, self.warmup, self.end, self.power = lr, total_num_update, warmup_updates, end_learning_rate, power
        self.lr = lr * (1.0 / warmup_updates if warmup_updates > 0 else 1.0)

    def step_begin_epoch(self, epoch):
        return self.lr

    def step_update(self, num_updates):
        if self.warmup > 0 and num_updates <= self.warmup:
            lr = num_updates / float(self.warmup) * self.base_lr
        elif num_updates >= self.total:
            lr = self.end
        else:
            pct = 1 - (num_updates - self.warmup) / (self.total - self.warmup)
            lr = (self.base_lr - self.end) * pct ** self.power + self.end
        self.lr = lr
        return lr
