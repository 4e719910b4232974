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

from . import lib as _lib
from . import ops as _ops


class Trainer:
    def __init__(self, model, criterion, lr_scheduler, adam_betas=(0.9, 0.98), adam_eps=1e-8, weight_decay=0.0,
                 clip_norm=2.0, process_group=None, use_cuda_graphs=False):
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
        # schedule scalars and the dropout seed live in device memory (written from a pinned staging buffer once
        # per update) so that a captured CUDA graph of the whole step can be replayed with fresh values
        self.use_cuda_graphs = use_cuda_graphs and dev.type == "cuda"
        self._hyper = torch.zeros(2, dtype=torch.float32, device=dev)      # [lr, step]
        self._seed = torch.zeros(1, dtype=torch.int64, device=dev)
        pin = dev.type == "cuda"
        self._hyper_host = torch.zeros(2, dtype=torch.float32, pin_memory=pin)
        self._seed_host = torch.zeros(1, dtype=torch.int64, pin_memory=pin)
        _ops.set_seed_tensor(self._seed)
        self._graphs, self._seen, self._pool = {}, {}, None

    def get_lr(self):
        return self.lr_scheduler.lr

    def _stage_scalars(self):
        lr = self.lr_scheduler.step_update(self.num_updates)
        self._hyper_host[0] = lr
        self._hyper_host[1] = float(self.num_updates + 1)
        self._seed_host[0] = (self.num_updates + 1) * 1000003
        self._hyper.copy_(self._hyper_host, non_blocking=True)
        self._seed.copy_(self._seed_host, non_blocking=True)

    def _fwd_bwd(self, samples):
        """Forward + hand-written backward of every micro-batch into the flat gradient buffer -- pure device work;
        this is the part captured in a CUDA graph per input shape."""
        model, flat = self.model, self.flat
        flat.zero_grad()
        tail = flat.tail
        for sample in samples:
            if not sample:
                continue
            loss, sample_size, log = self.criterion(model, sample)
            loss.backward()
            (model.sync_torch_grads_ if hasattr(model, "sync_torch_grads_") else model.encoder.sync_torch_grads_)()
            # logging scalars ride in the gradient buffer's tail
            tail[0] += sample_size
            tail[1] += log["ntokens"]
            tail[2] += log["nsentences"]
            tail[3] += log["loss"].float()

    def _reduce_and_update(self):
        """ONE collective (gradients + stats tail), then grad-norm + fused Adam reading lr/step/sample_size from device
        memory.  Kept outside the CUDA graph: the NCCL call stays an ordinary stream operation."""
        flat = self.flat
        if self.world > 1:
            dist.all_reduce(flat.g32, op=dist.ReduceOp.SUM, group=self.pg)
        _ops.sumsq(flat.grads, self._sumsq)
        _ops.adam_step(flat.p32, flat.m, flat.v, flat.grads, flat.p16, 0.0, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, 1, self._sumsq, denom_dev=flat.tail[0:1], clip_norm=self.clip_norm,
                       gnorm_out=self._gnorm, hyper_dev=self._hyper)

    def _step_body(self, samples):
        self._fwd_bwd(samples)
        self._reduce_and_update()

    @staticmethod
    def _signature(sample):
        ni = sample["net_input"]
        lens_cpu = ni.get("src_lengths_cpu")
        full = bool((lens_cpu == lens_cpu.max()).all()) if lens_cpu is not None else None
        sig = [("pads", full)]
        for k in sorted(ni):
            if torch.is_tensor(ni[k]) and ni[k].is_cuda:
                sig.append((k, tuple(ni[k].shape), str(ni[k].dtype)))
        sig.append(("target", tuple(sample["target"].shape)))
        return tuple(sig)

    def _graphed_step(self, sample):
        key = self._signature(sample)
        entry = self._graphs.get(key)
        if entry is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] < 2:  # first sight of this shape: plain eager step (also warms every kernel up)
                self._step_body([sample])
                return
            static = {"net_input": {k: (v.clone() if torch.is_tensor(v) and v.is_cuda else v)
                                    for k, v in sample["net_input"].items()},
                      "target": sample["target"].clone()}
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph, pool=self._pool):
                self._fwd_bwd([static])
            n_kernels = _lib.launch_count() - n0  # native kernels recorded in this graph
            _lib.load().esp_note_graph_replay(-n_kernels)  # capture itself executed nothing
            if self._pool is None:
                self._pool = graph.pool()
            entry = (graph, static, n_kernels)
            self._graphs[key] = entry
        graph, static, n_kernels = entry
        for k, v in sample["net_input"].items():
            if torch.is_tensor(v) and v.is_cuda:
                static["net_input"][k].copy_(v, non_blocking=True)
        static["target"].copy_(sample["target"], non_blocking=True)
        graph.replay()
        _lib.load().esp_note_graph_replay(n_kernels)
        self._reduce_and_update()

    def train_step(self, samples):
        """samples: list of micro-batches (update_freq entries); an empty dict is a dummy batch whose
        contribution is zero but which still joins the collective (trainer.py:1305-1313).

        With use_cuda_graphs, single-micro-batch updates are captured per input shape (second occurrence) and
        replayed afterwards: the ~1200 kernel launches of a step cost one graph launch on the host."""
        model = self.model
        model.train()
        model.set_num_updates(self.num_updates)
        self._stage_scalars()
        if self.use_cuda_graphs and len(samples) == 1 and samples[0] and "src_lengths_cpu" in samples[0]["net_input"]:
            self._graphed_step(samples[0])
        else:
            self._step_body(samples)
        self.num_updates += 1
        self.last_stats = self.flat.tail  # device tensor; read it lazily (e.g. every log_interval updates)
        return self.last_stats

    def stats(self):
        """Host copy of the last update's global stats (one sync; call at log intervals only)."""
        t = self.last_stats.detach().cpu().tolist()
        return {"sample_size": t[0], "ntokens": t[1], "nsentences": t[2], "loss": t[3], "gnorm": float(self._gnorm.item()),
                "lr": self.get_lr(), "num_updates": self.num_updates}
