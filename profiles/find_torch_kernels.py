"""Which Python lines launch the non-native (ATen / cuDNN) kernels of one eager training step?
torch.profiler with stacks; prints, per ATen op that launched CUDA kernels, the count, CUDA time and the innermost
repo frame.   python profiles/find_torch_kernels.py"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from espresso_b200 import lib  # noqa: E402
from espresso_b200.criterions import CtcLossCriterion  # noqa: E402
from espresso_b200.data.frontend import OnTheFlyFbank  # noqa: E402
from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel  # noqa: E402
from espresso_b200.optim import NoamLRScheduler  # noqa: E402
from espresso_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda:0")
lib.load()
torch.manual_seed(1)
cfg = SpeechTransformerConfig.from_dict(dict(dropout=0.1, attention_dropout=0.1, activation_dropout=0.1,
                                             layernorm_embedding=True, encoder=dict(bench.MODEL)))
model = SpeechTransformerEncoderModel.build_model(cfg, bench._Task()).finalize_(dev)
model.frontend = OnTheFlyFbank(np.full(80, 15.0), np.full(80, 4.0))
trainer = Trainer(model, CtcLossCriterion(bench._Task()), NoamLRScheduler(5.0, 25000, 512, 1e-6), adam_betas=(0.9, 0.98),
                  clip_norm=2.0, use_cuda_graphs=False)
b = bench.make_batches(1, 1, 0)[0]
d = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
sample = {"net_input": {"src_tokens": d["wave"], "src_lengths": d["n_samples"], "freq_masks": d["fm"], "time_masks": d["tm"],
                        "src_lengths_cpu": torch.from_numpy(b["n_samples"]).long()}, "target": d["target"]}
for _ in range(2):
    trainer.train_step([sample])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    trainer.train_step([sample])
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_stack_n=12):
    t = getattr(ev, "self_device_time_total", 0) or 0
    if t <= 0 or not ev.key.startswith("aten::"):
        continue
    frame = next((f for f in ev.stack if "/espresso_b200/" in f or "bench.py" in f), "?")
    rows.append((t, ev.count, ev.key, frame.split("/root/repo/")[-1][:110]))
for t, n, name, frame in sorted(rows, reverse=True)[:40]:
    print("%5d x %9.1f us  %-26s %s" % (n, t, name, frame))
