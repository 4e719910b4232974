"""Data-parallel training step, B200-native mirror of `Trainer.train_step` (fairseq/trainer.py:780-1097).

One process per GPU (torchrun / torch.distributed, backend nccl).  Per update:
  zero flat grads -> for each micro-batch: criterion forward + hand-written backward (kernels accumulate into
  the flat fp32 gradient buffer) -> ONE all-reduce over that buffer, whose 8-float tail carries
  [sample_size, ntokens, nsentences, loss] (replaces the reference's three collectives: stats float64
  all-reduce trainer.py:1411-1449, gradient all-reduce legacy_distributed_data_parallel.py:76-165 and the
  grad-norm consistency check trainer.py:1451-1488) -> esp_sumsq_f32 -> esp_adam_step, which reads
  1/sample_size and computes the clip coefficient on the device (trainer.py:918-953 semantics: grads are
  normalised by the global sample_size, clipped to clip_norm, Adam with fp32 master weights, bf16 params
  refreshed).  No host<->device synchronisation inside the step.
"""
import torch
import torch.distributed as dist

from . import ops as _ops


class Trainer:
    def __init__(self, model, criterion, lr_scheduler, adam_betas=(0.9, 0.98), adam_eps=1e-8, weight_decay=0.0,
                 clip_norm=2.0, process_group=None):
        self.model = model
        self.criterion = criterion
        self.lr_scheduler = lr_scheduler
        self.betas, self.eps, self.weight_decay, self.clip_norm = adam_betas, adam_eps, weight_decay, clip_norm
        self.flat = model.flat
        self.flat.init_master()
        dev = self.flat.p16.device
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.num_updates = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.last_stats = None

    def get_lr(self):
        return self.lr_scheduler.lr

    def train_step(self, samples):
        """samples: list of micro-batches (update_freq entries); an empty dict is a dummy batch whose
        contribution is zero but which still joins the collective (trainer.py:1305-1313)."""
        model, flat = self.model, self.flat
        model.train()
        model.set_num_updates(self.num_updates)
        flat.zero_grad()
        tail = flat.tail
        for sample in samples:
            if not sample:
                continue
            loss, sample_size, log = self.criterion(model, sample)
            loss.backward()
            model.encoder.sync_torch_grads_()
            # logging scalars ride in the gradient buffer's tail
            tail[0] += sample_size
            tail[1] += log["ntokens"]
            tail[2] += log["nsentences"]
            tail[3] += log["loss"].float()
        if self.world > 1:
            dist.all_reduce(flat.g32, op=dist.ReduceOp.SUM, group=self.pg)
        lr = self.lr_scheduler.step_update(self.num_updates)
        _ops.sumsq(flat.grads, self._sumsq)
        _ops.adam_step(flat.p32, flat.m, flat.v, flat.grads, flat.p16, lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.num_updates + 1, self._sumsq, denom_dev=tail[0:1],
                       clip_norm=self.clip_norm, gnorm_out=self._gnorm)
        self.num_updates += 1
        self.last_stats = tail  # device tensor; read it lazily (e.g. every log_interval updates)
        return tail

    def stats(self):
        """Host copy of the last update's global stats (one sync; call at log intervals only)."""
        t = self.last_stats.detach().cpu().tolist()
        return {"sample_size": t[0], "ntokens": t[1], "nsentences": t[2], "loss": t[3], "gnorm": float(self._gnorm.item()),
                "lr": self.get_lr(), "num_updates": self.num_updates}
