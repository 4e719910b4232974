"""Timeline (ns, relative) of CTA (0,0,0) of the fused attention backward kernel: where a query-tile iteration spends its time.
    python profiles/attn_bwd_timeline.py [B T]"""
import ctypes
import os
import sys

import torch

os.environ.setdefault("ESP_ATTN_BWD_TIMELINE", "1")
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
ctx, p, pd = ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, None, drop_p=0.1, seed=3)
dctx = (torch.randn(R, d, device=dev) * 0.5).bfloat16()
dqkv = torch.empty(R, 3 * d, device=dev, dtype=torch.bfloat16)
ldp = (2 * T - 1 + 7) // 8 * 8
for _ in range(3):
    ops.attn_fused_bwd(dctx, ctx, qu, v, p, pd, B, T, H, ldp, dqkv[:, d:2 * d], dqkv[:, 2 * d:], 0.1, 3)
buf = (ctypes.c_ulonglong * 128)()
L = lib.load()
L.esp_attn_bwd_timeline.argtypes = [ctypes.c_void_p]
L.esp_attn_bwd_timeline(buf)
t = list(buf)
t0 = t[0]
names = {0: "prologue done", 1: "pdl wait done", 2: "V loaded", 100: "accumulators final (warp 0)", 101: "CTA done"}
nq = (T + 127) // 128
for i in range(min(nq, 8)):
    names[4 + 4 * i] = "ctl  tile %d: stage loaded" % i
    names[5 + 4 * i] = "ctl  tile %d: dS(%d) seen" % (i, i - 1)
    names[6 + 4 * i] = "ctl  tile %d: stage(%d) free (MMAs)" % (i, i - 1)
    names[7 + 4 * i] = "ctl  tile %d: stage(%d) read (skew stores) -> load %d" % (i, i - 1, i + 1)
    names[64 + 4 * i] = "warp tile %d: dPd ready" % i
    names[65 + 4 * i] = "warp tile %d: dS written (arrive)" % i
    names[66 + 4 * i] = "warp tile %d: dS rows stored, pair synced" % i
    names[67 + 4 * i] = "warp tile %d: skewed rows stored" % i
ev = sorted((t[s] - t0, names[s]) for s in names if t[s] >= t0 and t[s] - t0 < 10 ** 9)
for ns, nm in ev:
    print("%8d ns  %s" % (ns, nm))
