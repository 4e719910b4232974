"""One GEMM shape, a few launches -- the target of `ncu --set full` captures (profiles/README.md).
    python profiles/gemm_one.py M N K tile_n [epilogue]     epilogue: plain | ffn1 (bias+SiLU+C2+dropout)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

M, N, K, tn = (int(a) for a in sys.argv[1:5])
epi = sys.argv[5] if len(sys.argv) > 5 else "plain"
dev = torch.device("cuda:0")
x = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
W = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
b = torch.randn(N, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
U = torch.empty_like(out)
for _ in range(3):
    if epi == "ffn1":
        ops.linear(x, W, b, act=ops.ACT_SILU, C2=U, drop_p=0.1, drop_mode=1, seed=1, out=out, tile_n=tn)
    else:
        ops.linear(x, W, out=out, tile_n=tn)
torch.cuda.synchronize()
