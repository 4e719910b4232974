"""GEMM micro-benchmark over the shapes of one Conformer layer (B=16 utterances x T'=406 frames, d=512).
Times 30 back-to-back launches per (shape, tile_n) with CUDA events.  Run on the B200 box:
    python profiles/gemm_microbench.py > gpurun_out/gemm_microbench.txt
"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, d, H, hd, ffn = 16, 406, 512, 8, 64, 2048
R = B * T
BF = torch.bfloat16


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
    return s.elapsed_time(e) / n * 1e3  # us


def report(name, flops, us):
    print("%-46s %8.1f us  %7.1f TFLOP/s" % (name, us, flops / us / 1e6))


def rnd(*s):
    return (torch.randn(*s, device=dev) * 0.1).to(BF)


x = rnd(R, d)
h = rnd(R, ffn)
for tn in (0, 256, 512):  # 0 = heuristic, 256 = single-CTA/multicast 128x256, 512 = cta_group::2 256x256 per CTA pair
    W1, b1 = rnd(ffn, d), rnd(ffn)
    U = torch.empty(R, ffn, device=dev, dtype=BF)
    out = torch.empty(R, ffn, device=dev, dtype=BF)
    report("fwd FFN1 plain            tile_n=%d" % tn, 2.0 * R * ffn * d, timeit(lambda: ops.linear(x, W1, out=out, tile_n=tn)))
    report("fwd FFN1 bias+silu+C2+drop tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.linear(x, W1, b1, act=ops.ACT_SILU, C2=U, drop_p=0.1, drop_mode=1, seed=1, out=out, tile_n=tn)))
    W2 = rnd(d, ffn)
    o2 = torch.empty(R, d, device=dev, dtype=BF)
    report("fwd FFN2 +residual        tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.linear(h, W2, None, R=x, ldr=d, alpha=0.5, out=o2, tile_n=tn)))
    Wq = rnd(3 * d, d)
    o3 = torch.empty(R, 3 * d, device=dev, dtype=BF)
    report("fwd QKV                   tile_n=%d" % tn, 2.0 * R * 3 * d * d, timeit(lambda: ops.linear(x, Wq, out=o3, tile_n=tn)))
    Wo = rnd(d, d)
    report("fwd out_proj (N=512)      tile_n=%d" % tn, 2.0 * R * d * d, timeit(lambda: ops.linear(x, Wo, out=o2, tile_n=tn)))
    # dgrad: dx[M,K] = dy[M,N] @ W[N,K]
    dU = rnd(R, ffn)
    report("dgrad FFN1 (M=R,N=512,K=2048) tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.gemm(dU, W1, o2, R, d, ffn, ffn, d, d, b_kmajor=False, tile_n=tn)))
    dZ = rnd(R, d)
    report("dgrad FFN2 +silu_bwd+drop tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.gemm(dZ, W2, out, R, ffn, d, d, ffn, ffn, b_kmajor=False, act=ops.ACT_SILU_BWD, aux=U, ld_aux=ffn,
                                   drop_p=0.1, drop_mode=2, seed=1, tile_n=tn)))
    # wgrad (accumulate, split-K)
    g1 = torch.zeros(ffn, d, device=dev)
    report("wgrad dW1 [2048,512] K=R  tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.gemm(dU, x, g1, ffn, d, R, ffn, d, d, a_kmajor=False, b_kmajor=False, accumulate=True, tile_n=tn)))
    g2 = torch.zeros(d, ffn, device=dev)
    report("wgrad dW2 [512,2048] K=R  tile_n=%d" % tn, 2.0 * R * ffn * d,
           timeit(lambda: ops.gemm(dZ, h, g2, d, ffn, R, d, ffn, ffn, a_kmajor=False, b_kmajor=False, accumulate=True, tile_n=tn)))
    g3 = torch.zeros(d, d, device=dev)
    report("wgrad dWo [512,512] K=R   tile_n=%d" % tn, 2.0 * R * d * d,
           timeit(lambda: ops.gemm(dZ, x, g3, d, d, R, d, d, d, a_kmajor=False, b_kmajor=False, accumulate=True, tile_n=tn)))
    print()

# attention-shaped batched GEMMs
q = rnd(R, d)
qkv = rnd(R, 3 * d)
ldt = (T + 7) // 8 * 8
S = torch.empty(H, B, T, ldt, device=dev, dtype=BF)
for tn in (0, 64, 128, 256):
    report("QK^T  [T,T,64] x %d heads   tile_n=%d" % (H * B, tn), 2.0 * H * B * T * T * hd,
           timeit(lambda: ops.gemm(q, qkv[:, d:2 * d], S, T, T, hd, d, 3 * d, ldt, nb1=H, nb2=B, sA=(hd, T * d), sB=(hd, T * 3 * d),
                                   sC=(B * T * ldt, T * ldt), tile_n=tn)))
ctx = torch.empty(R, d, device=dev, dtype=BF)
for tn in (0, 64):
    report("P V   [T,64,T] x %d heads   tile_n=%d" % (H * B, tn), 2.0 * H * B * T * T * hd,
           timeit(lambda: ops.gemm(S, qkv[:, 2 * d:], ctx, T, hd, T, ldt, 3 * d, d, b_kmajor=False, nb1=H, nb2=B,
                                   sA=(B * T * ldt, T * ldt), sB=(hd, T * 3 * d), sC=(hd, T * d), tile_n=tn)))
# reference points: cuBLAS via torch for the same FFN shapes
W1 = rnd(ffn, d)
report("torch.matmul FFN1 (cuBLAS reference point)", 2.0 * R * ffn * d, timeit(lambda: torch.matmul(x, W1.t())))
report("torch.matmul dgrad FFN1", 2.0 * R * ffn * d, timeit(lambda: torch.matmul(h, W1)))
report("torch.matmul wgrad dW1", 2.0 * R * ffn * d, timeit(lambda: torch.matmul(h.t(), x)))
