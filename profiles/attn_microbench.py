"""Rel-pos attention forward at the bench's typical shape: fused kernel (csrc/attn_fused.cu) vs the round-1 chain
(BD GEMM -> QK^T+skew GEMM -> softmax -> P V GEMM).  CUDA events over back-to-back launches.
    python profiles/attn_microbench.py [B T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 407)
H, hd = 8, 64
d = H * hd
R = B * T
torch.manual_seed(0)
qkv = (torch.randn(R, 3 * d, device=dev) * 0.5).bfloat16()
qu = (torch.randn(R, d, device=dev) * 0.3).bfloat16()
qv = (torch.randn(R, d, device=dev) * 0.3).bfloat16()
pos = (torch.randn(2 * T - 1, d, device=dev) * 0.5).bfloat16()
k, v = qkv[:, d:2 * d], qkv[:, 2 * d:]
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
lens[B // 2:] = int(0.8 * T)
r8 = lambda n: (n + 7) // 8 * 8  # noqa: E731
ldt, ldp = r8(T), r8(2 * T - 1)


def chain(drop):
    BD = torch.empty(H, B, T, ldp, device=dev, dtype=torch.bfloat16)
    ops.gemm(qv, pos, BD, T, 2 * T - 1, hd, d, d, ldp, nb1=H, nb2=B, sA=(hd, T * d), sB=(hd, 0), sC=(B * T * ldp, T * ldp))
    S = torch.empty(H, B, T, ldt, device=dev, dtype=torch.bfloat16)
    ops.gemm(qu, k, S, T, T, hd, d, 3 * d, ldt, nb1=H, nb2=B, sA=(hd, T * d), sB=(hd, T * 3 * d), sC=(B * T * ldt, T * ldt),
             R=BD, ldr=ldp, sR=(B * T * ldp, T * ldp), skew_r=T)
    Pr, Pd = ops.attn_softmax_fwd(S, T, lens, drop, 11)
    ctx = torch.empty(R, d, device=dev, dtype=torch.bfloat16)
    ops.gemm(Pd, v, ctx, T, hd, T, ldt, 3 * d, d, b_kmajor=False, nb1=H, nb2=B, sA=(B * T * ldt, T * ldt), sB=(hd, T * 3 * d),
             sC=(hd, T * d))
    return ctx


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


flops = 2.0 * B * H * hd * (T * T + T * (2 * T - 1) + T * T)
for drop in (0.0, 0.1):
    a = timeit(lambda: chain(drop))
    f = timeit(lambda: ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, lens, drop_p=drop, seed=11))
    g = timeit(lambda: ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, lens, drop_p=drop, seed=11, save_probs=False))
    print("B=%d T=%d dropout %.1f: round-1 chain %7.1f us | fused (saves P%s) %7.1f us = %5.1f TFLOP/s | fused, inference %7.1f us"
          % (B, T, drop, a, "+Pd" if drop else "", f, flops / f / 1e6, g))
c0 = chain(0.0)
c1, _, _ = ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, lens)
print("max |ctx_fused - ctx_chain| = %.3e (bf16 scores in the chain, fp32 in the fused kernel)" % (c0.float() - c1.float()).abs().max().item())

# ---- backward, score side: fused kernel (csrc/attn_fused_bwd.cu) vs the unfused chain it replaces ----
dctx = (torch.randn(R, d, device=dev) * 0.5).bfloat16()
dqkv = torch.empty(R, 3 * d, device=dev, dtype=torch.bfloat16)


def bwd_chain(ctx, p, pd, drop):
    dPd = torch.empty(H, B, T, ldt, device=dev, dtype=torch.bfloat16)
    ops.gemm(dctx, v, dPd, T, T, hd, d, 3 * d, ldt, nb1=H, nb2=B, sA=(hd, T * d), sB=(hd, T * 3 * d), sC=(B * T * ldt, T * ldt))
    ops.gemm(pd, dctx, dqkv[:, 2 * d:], T, hd, T, ldt, d, 3 * d, a_kmajor=False, b_kmajor=False, nb1=H, nb2=B,
             sA=(B * T * ldt, T * ldt), sB=(hd, T * d), sC=(hd, T * 3 * d))
    dS, dBD = ops.attn_softmax_bwd(p, dPd, T, ldp, drop, 11)
    ops.gemm(dS, qu, dqkv[:, d:2 * d], T, hd, T, ldt, d, 3 * d, a_kmajor=False, b_kmajor=False, nb1=H, nb2=B,
             sA=(B * T * ldt, T * ldt), sB=(hd, T * d), sC=(hd, T * 3 * d))
    return dS, dBD


for drop in (0.0, 0.1):
    ctx, p, pd = ops.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, lens, drop_p=drop, seed=11)
    a = timeit(lambda: bwd_chain(ctx, p, pd, drop))
    f = timeit(lambda: ops.attn_fused_bwd(dctx, ctx, qu, v, p, pd, B, T, H, ldp, dqkv[:, d:2 * d], dqkv[:, 2 * d:], drop, 11))
    print("backward (dPd, dS/dBD, dV, dK) B=%d T=%d dropout %.1f: unfused chain %7.1f us | fused %7.1f us" % (B, T, drop, a, f))
