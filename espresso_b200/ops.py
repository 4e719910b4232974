"""Thin torch-tensor wrappers over the C ABI (include/espresso_b200.h).

PyTorch is used only for device memory and the current CUDA stream; all arithmetic happens in the
hand-written sm_100a kernels.  Every wrapper requires CUDA tensors and raises otherwise.
"""
import ctypes as C

import torch

from . import lib as _lib
from .lib import ACT_NONE, ACT_RELU, ACT_RELU_BWD, ACT_SILU, ACT_SILU_BWD, EspGemm  # noqa: F401


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.EspressoB200Error("espresso_b200 ops need CUDA tensors (no CPU fallback)")


# ----------------------------------------------------------------------------------------------
# dense contraction
# ----------------------------------------------------------------------------------------------
def gemm(A, B, C_out, M, N, K, lda, ldb, ldc, *, a_kmajor=True, b_kmajor=True, nb1=1, nb2=1,
         sA=(0, 0), sB=(0, 0), sC=(0, 0), bias=None, act=ACT_NONE, aux=None, ld_aux=0, sAux=(0, 0),
         R=None, ldr=0, sR=(0, 0), alpha=1.0, beta=1.0, C2=None, drop_p=0.0, drop_mode=0, seed=0,
         skew_r=0, tile_n=0):
    """C = epilogue(op(A) @ op(B)^T); see EspGemm in include/espresso_b200.h.  Strides in elements."""
    _need_cuda(A, B, C_out, bias, aux, R, C2)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    g = EspGemm()
    g.A, g.B, g.C = A.data_ptr(), B.data_ptr(), C_out.data_ptr()
    g.C2 = C2.data_ptr() if C2 is not None else None
    g.bias = bias.data_ptr() if bias is not None else None
    g.aux = aux.data_ptr() if aux is not None else None
    g.R = R.data_ptr() if R is not None else None
    if bias is not None:
        assert bias.dtype == torch.bfloat16
    if aux is not None:
        assert aux.dtype == torch.bfloat16
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc, g.ld_aux, g.ldr = lda, ldb, ldc, ld_aux, ldr
    g.sA1, g.sA2 = sA
    g.sB1, g.sB2 = sB
    g.sC1, g.sC2 = sC
    g.sAux1, g.sAux2 = sAux
    g.sR1, g.sR2 = sR
    g.a_kmajor, g.b_kmajor = int(a_kmajor), int(b_kmajor)
    g.nb1, g.nb2 = nb1, nb2
    g.c_f32 = int(C_out.dtype == torch.float32)
    g.r_f32 = int(R is not None and R.dtype == torch.float32)
    g.act, g.drop_mode, g.skew_r, g.tile_n = act, drop_mode, skew_r, tile_n
    g.alpha, g.beta, g.drop_p = alpha, beta, drop_p
    g.seed = seed
    _lib.check(_lib.load().esp_gemm_bf16(C.byref(g), _stream()))
    return C_out


def linear(x, W, bias=None, *, act=ACT_NONE, out=None, out_dtype=torch.bfloat16, **kw):
    """y[M,N] = epi(x[M,K] @ W[N,K]^T + bias); x, W contiguous 2-D (row strides may exceed K)."""
    M, K = x.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype)
    return gemm(x, W, out, M, N, K, x.stride(0), W.stride(0), out.stride(0), bias=bias, act=act, **kw)


# ----------------------------------------------------------------------------------------------
# front end
# ----------------------------------------------------------------------------------------------
def num_frames(n_samples):
    """snip_edges frame count for 25 ms / 10 ms at 16 kHz (espresso/tools/utils.py:457-486)."""
    return torch.where(n_samples >= 400, 1 + (n_samples - 400) // 160, torch.zeros_like(n_samples))


def frontend_fbank(wave, n_samples, cmvn_mean=None, cmvn_std=None, freq_masks=None, time_masks=None,
                   t_max=None, out_dtype=torch.bfloat16, workspace=None):
    """Fused fbank -> CMVN -> SpecAugment.  wave [B, N] fp32/int16 (int16 value range), n_samples int32 [B].

    Returns (feats [B, t_max, 80], lengths int32 [B]).
    """
    _need_cuda(wave, n_samples, cmvn_mean, cmvn_std, freq_masks, time_masks)
    assert wave.dim() == 2 and wave.stride(1) == 1
    assert wave.dtype in (torch.float32, torch.int16)
    assert n_samples.dtype == torch.int32
    B = wave.shape[0]
    if t_max is None:
        n_max = wave.shape[1]
        t_max = 1 + (n_max - 400) // 160 if n_max >= 400 else 0
    out = torch.empty(B, t_max, 80, device=wave.device, dtype=out_dtype)
    lens = torch.empty(B, device=wave.device, dtype=torch.int32)
    if workspace is None:
        workspace = torch.zeros(max(1, B) * 16, device=wave.device, dtype=torch.uint8)
    nf = 0 if freq_masks is None else freq_masks.shape[1]
    nt = 0 if time_masks is None else time_masks.shape[1]
    for m in (freq_masks, time_masks):
        if m is not None:
            assert m.dtype == torch.int32 and m.is_contiguous() and m.shape[0] == B and m.shape[2] == 2
    for s in (cmvn_mean, cmvn_std):
        if s is not None:
            assert s.dtype == torch.float32 and s.numel() == 80
    _lib.check(_lib.load().esp_frontend_fbank(
        _ptr(wave), int(wave.dtype == torch.int16), wave.stride(0), _ptr(n_samples), B, _ptr(cmvn_mean),
        _ptr(cmvn_std), _ptr(freq_masks), nf, _ptr(time_masks), nt, _ptr(out),
        int(out_dtype == torch.float32), t_max, _ptr(lens), _ptr(workspace), _stream()))
    return out, lens


# ----------------------------------------------------------------------------------------------
# CTC
# ----------------------------------------------------------------------------------------------
def ctc_loss(logits, V, in_lens, targets, tgt_lens, blank, zero_infinity=True, grad_scale=1.0, want_grad=True):
    """logits bf16 [B, T, ld] (ld >= V, batch-major); returns (loss fp32 [B], grad bf16 like logits or None)."""
    _need_cuda(logits, in_lens, targets, tgt_lens)
    assert logits.dtype == torch.bfloat16 and logits.dim() == 3 and logits.stride(2) == 1
    assert in_lens.dtype == torch.int32 and tgt_lens.dtype == torch.int32 and targets.dtype == torch.int32
    B, T, _ = logits.shape
    u_max = targets.shape[1] if targets.dim() == 2 else 0
    L = _lib.load()
    ws = torch.empty(int(L.esp_ctc_workspace_bytes(B, T, u_max)), device=logits.device, dtype=torch.uint8)
    loss = torch.empty(B, device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits) if want_grad else None
    _lib.check(L.esp_ctc_loss(_ptr(logits), logits.stride(0), logits.stride(1), V, B, T, _ptr(in_lens),
                              _ptr(targets.contiguous()), u_max, _ptr(tgt_lens), blank, int(zero_infinity),
                              float(grad_scale), _ptr(loss), _ptr(grad), _ptr(ws), _stream()))
    return loss, grad
