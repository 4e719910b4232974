"""Forward / backward orchestration of the speech Transformer decoder (teacher forcing) over the sm_100a kernels.

Host-side mirror of (reference tree):
  espresso/models/transformer/speech_transformer_decoder.py:43-281
  fairseq/models/transformer/transformer_decoder.py:254-378       (embed -> layers -> layer_norm -> output proj)
  fairseq/modules/transformer_layer.py:384-533                     (pre-LN decoder block: causal self-attention,
                                                                    encoder attention, ReLU FFN)
  fairseq/modules/multihead_attention.py:639-917                   (plain scaled dot-product branch)
  fairseq/modules/sinusoidal_positional_embedding.py               (absolute positions, right-padded targets)
Activations are batch-major [B*U, d] bf16.  Like the encoder engine this only sequences kernel calls and owns the
saved-for-backward tensors; the backward is written by hand and accumulates into the flat fp32 gradient buffer.
"""
import math

import torch

from .. import ops as _ops
from ..lib import ACT_RELU, ACT_RELU_BWD
from .encoder_engine import LN_EPS, EncoderEngine, _r8


def sinusoidal_positions(U, d, padding_idx, device, dtype=torch.bfloat16):
    """Rows t = 0..U-1 of fairseq's sinusoidal table at index t + padding_idx + 1 (right-padded targets),
    computed in fp32 then cast to the model dtype."""
    half = d // 2
    f = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    pos = torch.arange(padding_idx + 1, padding_idx + 1 + U, dtype=torch.float)[:, None] * f[None, :]
    e = torch.cat([torch.sin(pos), torch.cos(pos)], dim=1)
    if d % 2 == 1:
        e = torch.cat([e, torch.zeros(U, 1)], dim=1)
    return e.to(device=device, dtype=dtype).contiguous()


class DecoderEngine(EncoderEngine):
    _FFN = ("final_layer_norm", "fc1", "fc2")

    def __init__(self, flat, prefix, cfg):
        """cfg keys: embed_dim, ffn_dim, heads, layers, vocab, pad, dropout, attention_dropout, activation_dropout,
        layernorm_embedding, share_input_output_embed, no_scale_embedding."""
        super().__init__(flat, prefix, dict(cfg, layer_type="decoder"))
        self._pos = {}
        # Scheduled sampling in the reference feeds a [B, 1] token tensor per step, so its sinusoidal position module
        # (incremental branch, fairseq/modules/sinusoidal_positional_embedding.py:78-87: pos = seq_len = 1) gives EVERY
        # step the embedding of the first position; the model sets this flag for such updates to reproduce it.
        self.constant_position = False

    def positions(self, U, device):
        key = (U, str(device), self.constant_position)
        if key not in self._pos:
            tab = sinusoidal_positions(U, self.d, self.cfg["pad"], device)
            self._pos[key] = tab[:1].expand(U, -1).contiguous() if self.constant_position else tab
        return self._pos[key]

    def _proj(self, lp, which, names):
        ws = [self.pre + lp + "%s.%s_proj.weight" % (which, c) for c in names]
        bs = [self.pre + lp + "%s.%s_proj.bias" % (which, c) for c in names]
        n, d = len(names), self.d
        return (self.flat.span(self.flat.p16, ws, (n * d, d)), self.flat.span(self.flat.p16, bs, (n * d,)),
                self.flat.span(self.flat.g32, ws, (n * d, d)), self.flat.span(self.flat.g32, bs, (n * d,)))

    # ---- scaled dot-product attention core shared by self- and cross-attention -------------------
    def _attend(self, q, ldq, k, v, ldkv, B, Tq, Tk, lens, causal, li, op):
        """q [B*Tq, .] (row stride ldq), k/v [B*Tk, .] (row stride ldkv) -> ctx [B*Tq, d], saved (Pr, Pd)."""
        d, H, hd = self.d, self.H, self.hd
        ld = _r8(Tk)
        S = torch.empty(H, B, Tq, ld, device=q.device, dtype=torch.bfloat16)
        _ops.gemm(q, k, S, Tq, Tk, hd, ldq, ldkv, ld, nb1=H, nb2=B, sA=(hd, Tq * ldq), sB=(hd, Tk * ldkv),
                  sC=(B * Tq * ld, Tq * ld), alpha=self.scaling)
        Pr, Pd = _ops.attn_softmax_fwd(S, Tk, lens, self._drop("attention_dropout"), self._seed(li, op), causal=causal)
        ctx = torch.empty(B * Tq, d, device=q.device, dtype=torch.bfloat16)
        _ops.gemm(Pd, v, ctx, Tq, hd, Tk, ld, ldkv, d, b_kmajor=False, nb1=H, nb2=B, sA=(B * Tq * ld, Tq * ld),
                  sB=(hd, Tk * ldkv), sC=(hd, Tq * d))
        return ctx, (Pr, Pd)

    def _attend_bwd(self, dctx, saved, q, ldq, k, v, ldkv, dq, lddq, dk, dv, lddkv, B, Tq, Tk, li, op):
        """Writes dq [B*Tq, d] (row stride lddq) and dk/dv [B*Tk, d] (row stride lddkv)."""
        d, H, hd = self.d, self.H, self.hd
        Pr, Pd = saved
        ld = _r8(Tk)
        dPd = torch.empty(H, B, Tq, ld, device=dctx.device, dtype=torch.bfloat16)
        _ops.gemm(dctx, v, dPd, Tq, Tk, hd, d, ldkv, ld, nb1=H, nb2=B, sA=(hd, Tq * d), sB=(hd, Tk * ldkv),
                  sC=(B * Tq * ld, Tq * ld))
        # dV = Pd^T dctx
        _ops.gemm(Pd, dctx, dv, Tk, hd, Tq, ld, d, lddkv, a_kmajor=False, b_kmajor=False, nb1=H, nb2=B,
                  sA=(B * Tq * ld, Tq * ld), sB=(hd, Tq * d), sC=(hd, Tk * lddkv))
        dS, _ = _ops.attn_softmax_bwd(Pr, dPd, Tk, 0, self._drop("attention_dropout"), self._seed(li, op), want_dbd=False)
        # dq = scaling * dS k ; dk = scaling * dS^T q
        _ops.gemm(dS, k, dq, Tq, hd, Tk, ld, ldkv, lddq, b_kmajor=False, nb1=H, nb2=B, sA=(B * Tq * ld, Tq * ld),
                  sB=(hd, Tk * ldkv), sC=(hd, Tq * lddq), alpha=self.scaling)
        _ops.gemm(dS, q, dk, Tk, hd, Tq, ld, ldq, lddkv, a_kmajor=False, b_kmajor=False, nb1=H, nb2=B,
                  sA=(B * Tq * ld, Tq * ld), sB=(hd, Tq * ldq), sC=(hd, Tk * lddkv), alpha=self.scaling)

    # ---- causal self-attention -------------------------------------------------------------------
    def self_attn_fwd(self, x, lp, B, U, tgt_lens, li):
        d = self.d
        ln, mean, rstd = _ops.layer_norm_fwd(x, self.P(lp + "self_attn_layer_norm.weight"),
                                             self.P(lp + "self_attn_layer_norm.bias"), LN_EPS)
        W, b, _, _ = self._proj(lp, "self_attn", "qkv")
        qkv = _ops.linear(ln, W, b)
        ctx, att = self._attend(qkv[:, :d], 3 * d, qkv[:, d:2 * d], qkv[:, 2 * d:], 3 * d, B, U, U, tgt_lens, True, li, 30)
        y = _ops.linear(ctx, self.P(lp + "self_attn.out_proj.weight"), self.P(lp + "self_attn.out_proj.bias"),
                        drop_p=self._drop("dropout"), drop_mode=1, seed=self._seed(li, 31), R=x, ldr=x.stride(0), beta=1.0)
        return y, (x, mean, rstd, ln, qkv, att, ctx)

    def self_attn_bwd(self, dy, saved, lp, B, U, li):
        d = self.d
        x, mean, rstd, ln, qkv, att, ctx = saved
        dO = _ops.dropout(dy, self._drop("dropout"), self._seed(li, 31))
        self._wgrad(dO, ctx, self.G(lp + "self_attn.out_proj.weight"))
        _ops.colsum(dO, self.G(lp + "self_attn.out_proj.bias"))
        dctx = self._dgrad(dO, self.P(lp + "self_attn.out_proj.weight"))
        dqkv = torch.empty_like(qkv)
        self._attend_bwd(dctx, att, qkv[:, :d], 3 * d, qkv[:, d:2 * d], qkv[:, 2 * d:], 3 * d, dqkv[:, :d], 3 * d,
                         dqkv[:, d:2 * d], dqkv[:, 2 * d:], 3 * d, B, U, U, li, 30)
        W, _, gW, gb = self._proj(lp, "self_attn", "qkv")
        self._wgrad(dqkv, ln, gW)
        _ops.colsum(dqkv, gb)
        dln = self._dgrad(dqkv, W)
        return _ops.layer_norm_bwd(dln, x, mean, rstd, self.P(lp + "self_attn_layer_norm.weight"),
                                   self.G(lp + "self_attn_layer_norm.weight"), self.G(lp + "self_attn_layer_norm.bias"), dres=dy)

    # ---- encoder (cross) attention -----------------------------------------------------------------
    def cross_attn_fwd(self, x, enc, lp, B, U, Tk, enc_lens, li):
        d = self.d
        ln, mean, rstd = _ops.layer_norm_fwd(x, self.P(lp + "encoder_attn_layer_norm.weight"),
                                             self.P(lp + "encoder_attn_layer_norm.bias"), LN_EPS)
        q = _ops.linear(ln, self.P(lp + "encoder_attn.q_proj.weight"), self.P(lp + "encoder_attn.q_proj.bias"))
        Wkv, bkv, _, _ = self._proj(lp, "encoder_attn", "kv")
        kv = _ops.linear(enc, Wkv, bkv)  # [B*Tk, 2d]
        ctx, att = self._attend(q, d, kv[:, :d], kv[:, d:], 2 * d, B, U, Tk, enc_lens, False, li, 32)
        y = _ops.linear(ctx, self.P(lp + "encoder_attn.out_proj.weight"), self.P(lp + "encoder_attn.out_proj.bias"),
                        drop_p=self._drop("dropout"), drop_mode=1, seed=self._seed(li, 33), R=x, ldr=x.stride(0), beta=1.0)
        return y, (x, mean, rstd, ln, q, kv, att, ctx)

    def cross_attn_bwd(self, dy, saved, enc, denc, lp, B, U, Tk, li):
        d = self.d
        x, mean, rstd, ln, q, kv, att, ctx = saved
        dO = _ops.dropout(dy, self._drop("dropout"), self._seed(li, 33))
        self._wgrad(dO, ctx, self.G(lp + "encoder_attn.out_proj.weight"))
        _ops.colsum(dO, self.G(lp + "encoder_attn.out_proj.bias"))
        dctx = self._dgrad(dO, self.P(lp + "encoder_attn.out_proj.weight"))
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        self._attend_bwd(dctx, att, q, d, kv[:, :d], kv[:, d:], 2 * d, dq, d, dkv[:, :d], dkv[:, d:], 2 * d, B, U, Tk, li, 32)
        Wkv, _, gWkv, gbkv = self._proj(lp, "encoder_attn", "kv")
        self._wgrad(dkv, enc, gWkv)
        _ops.colsum(dkv, gbkv)
        # gradient w.r.t. the encoder output accumulates over decoder layers: denc += dkv @ Wkv
        M, N = dkv.shape
        _ops.gemm(dkv, Wkv, denc, M, d, N, dkv.stride(0), Wkv.stride(0), d, b_kmajor=False, R=denc, ldr=d, beta=1.0)
        self._wgrad(dq, ln, self.G(lp + "encoder_attn.q_proj.weight"))
        _ops.colsum(dq, self.G(lp + "encoder_attn.q_proj.bias"))
        dln = self._dgrad(dq, self.P(lp + "encoder_attn.q_proj.weight"))
        return _ops.layer_norm_bwd(dln, x, mean, rstd, self.P(lp + "encoder_attn_layer_norm.weight"),
                                   self.G(lp + "encoder_attn_layer_norm.weight"), self.G(lp + "encoder_attn_layer_norm.bias"),
                                   dres=dy)

    # ---- whole decoder ------------------------------------------------------------------------------
    def out_weight(self):
        return self.P("embed_tokens.weight") if self.cfg.get("share_input_output_embed", False) else self.P("output_projection.weight")

    def out_grad(self):
        return self.G("embed_tokens.weight") if self.cfg.get("share_input_output_embed", False) else self.G("output_projection.weight")

    def forward(self, tokens, enc, enc_lens, tgt_lens, save=True):
        """tokens int32 [B, U] (prev_output_tokens, right padded); enc bf16 [B, Tk, d]; enc_lens int32 [B] or None
        (no encoder padding); tgt_lens int32 [B] or None (no target padding).  Returns logits [B, U, ldV]."""
        cfg = self.cfg
        B, U = tokens.shape
        Tk = enc.shape[1]
        d, V = self.d, cfg["vocab"]
        scale = 1.0 if cfg.get("no_scale_embedding", False) else math.sqrt(d)
        tok = tokens.reshape(-1).contiguous()
        pdrop = self._drop("dropout")
        lne = cfg.get("layernorm_embedding", False)
        x = _ops.embed_fwd(tok, self.P("embed_tokens.weight"), self.positions(U, enc.device), U, scale, cfg["pad"],
                           0.0 if lne else pdrop, self._seed(998, 0))
        emb = None
        if lne:
            emb = x
            x, m0, r0 = _ops.layer_norm_fwd(emb, self.P("layernorm_embedding.weight"), self.P("layernorm_embedding.bias"), LN_EPS,
                                            drop_p=pdrop, seed=self._seed(998, 0))
            emb = (emb, m0, r0)
        enc2 = enc.reshape(B * Tk, d)
        states = []
        for li in range(cfg["layers"]):
            lp = "layers.%d." % li
            st = {}
            x, st["self"] = self.self_attn_fwd(x, lp, B, U, tgt_lens, li)
            x, st["cross"] = self.cross_attn_fwd(x, enc2, lp, B, U, Tk, enc_lens, li)
            x, st["ffn"] = self.ffn_fwd(x, lp, self._FFN, ACT_RELU, 1.0, li, 34)
            states.append(st)
        xf = x
        x, mf, rf = _ops.layer_norm_fwd(xf, self.P("layer_norm.weight"), self.P("layer_norm.bias"), LN_EPS)
        ldV = _r8(V)
        R = B * U
        logits = torch.zeros(R, ldV, device=x.device, dtype=torch.bfloat16) if ldV != V else \
            torch.empty(R, ldV, device=x.device, dtype=torch.bfloat16)
        W = self.out_weight()
        _ops.gemm(x, W, logits, R, V, d, d, W.stride(0), ldV)
        if save:
            self.saved = dict(B=B, U=U, Tk=Tk, tok=tok, emb=emb, enc2=enc2, states=states, fin=(xf, mf, rf), xL=x, scale=scale)
        return logits.view(B, U, ldV)

    def backward(self, dlogits):
        """dlogits [B, U, ldV] -> d(enc) [B, Tk, d]; parameter gradients go to the flat buffer."""
        s = self.saved
        cfg = self.cfg
        B, U, Tk, d, V = s["B"], s["U"], s["Tk"], self.d, cfg["vocab"]
        R = B * U
        ldV = _r8(V)
        dlog = dlogits.reshape(R, ldV)
        dl = dlog[:, :V] if ldV != V else dlog
        self._wgrad(dl, s["xL"], self.out_grad())
        dx = self._dgrad(dl, self.out_weight())
        xf, mf, rf = s["fin"]
        dx = _ops.layer_norm_bwd(dx, xf, mf, rf, self.P("layer_norm.weight"), self.G("layer_norm.weight"), self.G("layer_norm.bias"))
        denc = torch.zeros(B * Tk, d, device=dx.device, dtype=torch.bfloat16)
        for li in reversed(range(cfg["layers"])):
            lp = "layers.%d." % li
            st = s["states"][li]
            dx = self.ffn_bwd(dx, st["ffn"], lp, self._FFN, ACT_RELU_BWD, 1.0, li, 34)
            dx = self.cross_attn_bwd(dx, st["cross"], s["enc2"], denc, lp, B, U, Tk, li)
            dx = self.self_attn_bwd(dx, st["self"], lp, B, U, li)
        pdrop = self._drop("dropout")
        if s["emb"] is not None:
            e, m0, r0 = s["emb"]
            dx = _ops.layer_norm_bwd(dx, e, m0, r0, self.P("layernorm_embedding.weight"), self.G("layernorm_embedding.weight"),
                                     self.G("layernorm_embedding.bias"), drop_p=pdrop, seed=self._seed(998, 0))
            pdrop = 0.0
        _ops.embed_bwd(s["tok"], dx, self.G("embed_tokens.weight"), s["scale"], cfg["pad"], pdrop, self._seed(998, 0))
        self.saved = None
        return denc.view(B, Tk, d)


class IncrementalDecoder:
    """One-token-per-hypothesis decoder step for beam search (eval only), over the same flat parameters as
    DecoderEngine.  Mirrors the incremental branch of the reference decoder
    (fairseq/models/transformer/transformer_decoder.py:300-305 `prev_output_tokens[:, -1:]`,
    fairseq/modules/transformer_layer.py:384-533 with incremental_state, multihead_attention.py:639-760).

    State: per layer a K/V cache [T_max, N, 2d] written in place by the K/V projection GEMM, an ancestry table
    [T_max, N] (see csrc/decode_attn.cu) and the encoder-attention K/V [bsz, Tk, 2d] computed once per sentence
    (the reference's beamable encoder attention, multihead_attention.py:661-669)."""

    def __init__(self, engine: DecoderEngine):
        self.e = engine

    def init_state(self, enc, enc_lens, bsz, beam, t_max, reuse=None):
        """enc bf16 [bsz, Tk, d] or None (language model: no encoder attention).

        `reuse`: a state returned earlier for the SAME shapes -- its buffers are re-initialised in place (KV caches keep
        their addresses, ancestry tables are zeroed and put back in their initial roles, the encoder-attention K/V are
        recomputed into the same tensors), which is what lets a search step be replayed from a CUDA graph."""
        e = self.e
        d, Lr = e.d, e.cfg["layers"]
        N = bsz * beam
        dev = e.flat.p16.device
        Tk = enc.shape[1] if enc is not None else 0
        sig = (N, beam, bsz, t_max, Tk, enc_lens is not None)
        if reuse is not None and reuse.get("sig") == sig:
            st = reuse
            st["anc"], st["anc_alt"] = st["anc0"], st["anc1"]
            st["anc"].zero_()
            st["anc_alt"].zero_()
            if enc_lens is not None:
                st["enc_lens"].copy_(enc_lens)
        else:
            st = dict(N=N, beam=beam, bsz=bsz, t_max=t_max, sig=sig)
            st["enc_lens"] = enc_lens.clone() if enc_lens is not None else None
            st["kv"] = [torch.empty(t_max, N, 2 * d, device=dev, dtype=torch.bfloat16) for _ in range(Lr)]
            st["anc0"] = torch.zeros(t_max, N, device=dev, dtype=torch.int32)
            st["anc1"] = torch.zeros_like(st["anc0"])
            st["anc"], st["anc_alt"] = st["anc0"], st["anc1"]
            st["cross"] = [torch.empty(bsz, Tk, 2 * d, device=dev, dtype=torch.bfloat16) for _ in range(Lr)] if enc is not None else None
        if enc is not None:
            enc2 = enc.reshape(bsz * Tk, d)
            for li in range(Lr):
                Wkv, bkv, _, _ = e._proj("layers.%d." % li, "encoder_attn", "kv")
                _ops.linear(enc2, Wkv, bkv, out=st["cross"][li].view(bsz * Tk, 2 * d))
        return st

    @staticmethod
    def advance_without_compute(st):
        """Host-side effect of step() when its device work is replayed from a CUDA graph: the ancestry tables swap."""
        st["anc"], st["anc_alt"] = st["anc_alt"], st["anc"]

    def step(self, step, tokens, st, new_order):
        """tokens int32 [N, L] (column `step` is the newest token) -> logits bf16 [N, ldV]."""
        e = self.e
        cfg = e.cfg
        d, H, V = e.d, e.H, cfg["vocab"]
        N = st["N"]
        _ops.decode_update_ancestry(st["anc"], st["anc_alt"], new_order if step > 0 else None, step)
        st["anc"], st["anc_alt"] = st["anc_alt"], st["anc"]
        scale = 1.0 if cfg.get("no_scale_embedding", False) else math.sqrt(d)
        tok = tokens[:, step].contiguous()
        pos = e.positions(st["t_max"], tokens.device)[step: step + 1]
        x = _ops.embed_fwd(tok, e.P("embed_tokens.weight"), pos, 1, scale, cfg["pad"])
        if cfg.get("layernorm_embedding", False):
            x, _, _ = _ops.layer_norm_fwd(x, e.P("layernorm_embedding.weight"), e.P("layernorm_embedding.bias"), LN_EPS)
        for li in range(cfg["layers"]):
            lp = "layers.%d." % li
            ln, _, _ = _ops.layer_norm_fwd(x, e.P(lp + "self_attn_layer_norm.weight"), e.P(lp + "self_attn_layer_norm.bias"), LN_EPS)
            W, b, _, _ = e._proj(lp, "self_attn", "qkv")
            q = _ops.linear(ln, W[:d], b[:d])
            _ops.linear(ln, W[d:], b[d:], out=st["kv"][li][step])  # K | V of this step, written into the cache
            ctx = _ops.decode_self_attn(q, st["kv"][li], st["anc"], step + 1, H, e.scaling)
            x = _ops.linear(ctx, e.P(lp + "self_attn.out_proj.weight"), e.P(lp + "self_attn.out_proj.bias"), R=x, ldr=d, beta=1.0)
            if st["cross"] is not None:
                ln, _, _ = _ops.layer_norm_fwd(x, e.P(lp + "encoder_attn_layer_norm.weight"), e.P(lp + "encoder_attn_layer_norm.bias"),
                                               LN_EPS)
                q = _ops.linear(ln, e.P(lp + "encoder_attn.q_proj.weight"), e.P(lp + "encoder_attn.q_proj.bias"))
                ctx = _ops.decode_cross_attn(q, st["cross"][li], st["enc_lens"], st["beam"], H, e.scaling)
                x = _ops.linear(ctx, e.P(lp + "encoder_attn.out_proj.weight"), e.P(lp + "encoder_attn.out_proj.bias"), R=x, ldr=d,
                                beta=1.0)
            ln, _, _ = _ops.layer_norm_fwd(x, e.P(lp + "final_layer_norm.weight"), e.P(lp + "final_layer_norm.bias"), LN_EPS)
            hh = _ops.linear(ln, e.P(lp + "fc1.weight"), e.P(lp + "fc1.bias"), act=ACT_RELU)
            x = _ops.linear(hh, e.P(lp + "fc2.weight"), e.P(lp + "fc2.bias"), R=x, ldr=d, beta=1.0)
        x, _, _ = _ops.layer_norm_fwd(x, e.P("layer_norm.weight"), e.P("layer_norm.bias"), LN_EPS)
        ldV = _r8(V)
        logits = torch.empty(N, ldV, device=x.device, dtype=torch.bfloat16)
        Wo = e.out_weight()
        _ops.gemm(x, Wo, logits, N, V, d, d, Wo.stride(0), ldV)
        return logits
