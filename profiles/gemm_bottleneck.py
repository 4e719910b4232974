"""Bottleneck experiment for the tcgen05 GEMM (ESP_GEMM_DEBUG switches parts of the kernel off; results are WRONG by
design, only the time matters): FFN1-shaped GEMM [6512 x 2048 x 512] and a long-K wgrad shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


M, N, K = 6512, 2048, 512
x = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
W = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
names = {0: "full kernel", 1: "no epilogue stores", 8: "no epilogue at all (hand TMEM back)", 2: "no MMA", 4: "no TMA loads",
         14: "barrier skeleton only", 12: "MMA only (no TMA, no epilogue)",
         10: "TMA only (no MMA, no epilogue)"}
for tn in (256, 512):
    for dbg, nm in names.items():
        os.environ["ESP_GEMM_DEBUG"] = str(dbg)
        us = timeit(lambda: ops.linear(x, W, out=out, tile_n=tn))
        print("tile_n=%d  %-40s %7.1f us" % (tn, nm, us))
    print()
os.environ["ESP_GEMM_DEBUG"] = "0"
