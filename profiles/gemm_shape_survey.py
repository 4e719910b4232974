"""Which GEMM shapes does one training step launch, and how fast is each?

Records every ops.gemm call of one eager step of the bench workload (shape, layout, epilogue flags), then times
each distinct configuration alone (10 back-to-back launches in a CUDA graph, CUDA events) and prints a table sorted
by the time the configuration contributes to a step.

    python profiles/gemm_shape_survey.py [--layers N]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from espresso_b200 import lib, ops  # noqa: E402
from espresso_b200.criterions import CtcLossCriterion  # noqa: E402
from espresso_b200.data.frontend import OnTheFlyFbank  # noqa: E402
from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel  # noqa: E402
from espresso_b200.optim import NoamLRScheduler  # noqa: E402
from espresso_b200.trainer import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib.load()
    torch.manual_seed(1)
    enc = dict(bench.MODEL)
    if args.layers:
        enc["layers"] = args.layers
    cfg = SpeechTransformerConfig.from_dict(dict(dropout=0.1, attention_dropout=0.1, activation_dropout=0.1,
                                                 layernorm_embedding=True, encoder=enc))
    model = SpeechTransformerEncoderModel.build_model(cfg, bench._Task()).finalize_(dev)
    model.frontend = OnTheFlyFbank(np.full(80, 15.0), np.full(80, 4.0))
    trainer = Trainer(model, CtcLossCriterion(bench._Task()), NoamLRScheduler(5.0, 25000, 512, 1e-6),
                      adam_betas=(0.9, 0.98), clip_norm=2.0, use_cuda_graphs=False)
    host = bench.make_batches(2, 1, 0)

    def sample(b):
        d = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        return {"net_input": {"src_tokens": d["wave"], "src_lengths": d["n_samples"], "freq_masks": d["fm"],
                              "time_masks": d["tm"], "src_lengths_cpu": torch.from_numpy(b["n_samples"]).long()},
                "target": d["target"]}

    trainer.train_step([sample(host[0])])  # warm
    calls = {}
    real = ops.gemm

    def rec(A, B, C_out, M, N, K, lda, ldb, ldc, **kw):
        key = (M, N, K, kw.get("nb1", 1), kw.get("nb2", 1), int(kw.get("a_kmajor", True)), int(kw.get("b_kmajor", True)),
               kw.get("act", 0), int(kw.get("drop_p", 0.0) > 0), int(kw.get("bias") is not None),
               int(kw.get("aux") is not None), int(kw.get("R") is not None), int(kw.get("C2") is not None),
               int(kw.get("accumulate", False)), int(C_out.dtype == torch.float32), kw.get("skew_r", 0))
        if key not in calls:
            calls[key] = [0, (A, B, C_out, M, N, K, lda, ldb, ldc, dict(kw))]
        calls[key][0] += 1
        return real(A, B, C_out, M, N, K, lda, ldb, ldc, **kw)

    ops.gemm = rec
    trainer.train_step([sample(host[1])])
    ops.gemm = real
    torch.cuda.synchronize()

    rows = []
    for key, (cnt, (A, B, C_out, M, N, K, lda, ldb, ldc, kw)) in calls.items():
        scratch = C_out  # strided views must keep their storage; accumulating into the live buffer is harmless here
        kw2 = dict(kw)

        def run():
            real(A, B, scratch, M, N, K, lda, ldb, ldc, **kw2)

        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 30
        fl = 2.0 * M * N * K * key[3] * key[4]
        rows.append((cnt * us, cnt, us, fl / us / 1e6, key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("total GEMM time per step (isolated launches, warm L2): %.2f ms over %d launches, %d distinct configs"
          % (tot / 1e3, sum(r[1] for r in rows), len(rows)))
    print("| share | n | us | TFLOP/s | M | N | K | nb | aK bK | act drop bias aux R C2 acc f32 skew |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for t, cnt, us, tf, key in rows:
        M, N, K, nb1, nb2, ak, bk, *flags = key
        print("| %4.1f%% | %3d | %6.1f | %6.1f | %d | %d | %d | %dx%d | %d %d | %s |"
              % (100 * t / tot, cnt, us, tf, M, N, K, nb1, nb2, ak, bk, " ".join(str(f) for f in flags)))


if __name__ == "__main__":
    main()
