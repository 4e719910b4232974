"""One eager training step of the bench workload between cudaProfilerStart/Stop -- the target of the ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/one_step.py [batch index]
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/r02_launches_eager_stepN.md"""
import os
import sys

import numpy as np
import torch

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
cfg = SpeechTransformerConfig.from_dict(dict(dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, layernorm_embedding=True,
                                             encoder=dict(bench.MODEL)))
model = SpeechTransformerEncoderModel.build_model(cfg, bench._Task()).finalize_(dev)
model.frontend = OnTheFlyFbank(np.full(80, 15.0), np.full(80, 4.0))
trainer = Trainer(model, CtcLossCriterion(bench._Task()), NoamLRScheduler(5.0, 25000, 512, 1e-6), adam_betas=(0.9, 0.98), clip_norm=2.0,
                  use_cuda_graphs=False)
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = bench.make_batches(which + 1)[which]
d = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
sample = {"net_input": {"src_tokens": d["wave"], "src_lengths": d["n_samples"], "freq_masks": d["fm"], "time_masks": d["tm"],
                        "src_lengths_cpu": torch.from_numpy(b["n_samples"]).long()}, "target": d["target"]}
for _ in range(2):
    trainer.train_step([sample])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trainer.train_step([sample])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one step done: B=%d frames=%d audio %.1f s" % (len(b["frames"]), int(b["frames"].sum()), b["audio_s"]))
