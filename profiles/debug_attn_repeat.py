"""Run-to-run determinism of the fused attention kernel (a race would show up as differing bits)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(3)
for (B, T, H, lens) in [(3, 250, 8, [250, 160, 77]), (16, 407, 8, None), (6, 875, 8, [875, 800, 700, 600, 300, 100])]:
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.7).bfloat16()
    qu = (torch.randn(B * T, d, device=dev) * 0.3).bfloat16()
    qv = (torch.randn(B * T, d, device=dev) * 0.3).bfloat16()
    pos = (torch.randn(2 * T - 1, d, device=dev) * 0.7).bfloat16()
    lt = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=dev)
    ref = None
    bad = 0
    for it in range(40):
        ctx, p, _ = ops.attn_fused_fwd(qu, qv, qkv[:, d:2 * d], qkv[:, 2 * d:], pos, B, T, H, lt)
        if ref is None:
            ref = (ctx.clone(), p.clone())
        else:
            bad += int(not (torch.equal(ctx, ref[0]) and torch.equal(p, ref[1])))
    print("B=%d T=%d: %d of 39 repeats differ from the first run" % (B, T, bad))
