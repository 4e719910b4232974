"""Timeline (ns, relative) of softmax warp 0 of CTA (0,0,0) of the fused attention forward: per (key tile, pass) iteration --
wait for the logits, logits in registers, iteration done (pass 2 also writes P / P_drop and the A operand).
    python profiles/attn_fwd_timeline.py [B T]"""
import ctypes
import os
import sys

import torch

os.environ.setdefault("ESP_ATTN_FWD_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import lib, ops  # noqa: E402

dev = torch.device("cuda:0")
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 407)
H, hd = 8, 64
d = H * hd
R = B * T
qkv = (torch.randn(R, 3 * d, device=dev) * 0.5).bfloat16()
qu = (torch.randn(R, d, device=dev) * 0.3).bfloat16()
qv = (torch.randn(R, d, device=dev) * 0.3).bfloat16()
pos = (torch.randn(2 * T - 1, d, device=dev) * 0.5).bfloat16()
k, v = qkv[:, d:2 * d], qkv[:, 2 * d:]
for _ in range(3):
    ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, None, drop_p=0.1, seed=3)
buf = (ctypes.c_ulonglong * 128)()
L = lib.load()
L.esp_attn_fwd_timeline.argtypes = [ctypes.c_void_p]
L.esp_attn_fwd_timeline(buf)
t = list(buf)
nkt = (T + 127) // 128
print("start -> first logits wait: %d ns" % (t[1] - t[0]))
for it in range(2 * nkt):
    a, b, c = t[1 + 3 * it], t[2 + 3 * it], t[3 + 3 * it]
    nxt = t[1 + 3 * (it + 1)] if it + 1 < 2 * nkt else c
    print("iter %d (pass %d, key tile %d): logits ready at %6d | staging + logits -> regs %5d ns | exp / stores %5d ns | wait for next %5d ns"
          % (it, 1 + it // nkt, it % nkt, a - t[0], b - a, c - b, nxt - c))
