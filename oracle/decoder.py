"""Oracle: Transformer decoder (teacher forcing) + label-smoothed CE, functional PyTorch on a state_dict with
the reference's parameter names.  TEST INFRASTRUCTURE ONLY.  Restates:
  espresso/models/transformer/speech_transformer_decoder.py:43-281 (SpeechTransformerDecoderBase)
  fairseq/models/transformer/transformer_decoder.py:254-378 (extract_features_scriptable, output_layer)
  fairseq/modules/transformer_layer.py:384-533 (TransformerDecoderLayerBase.forward, normalize_before=True)
  fairseq/modules/multihead_attention.py:639-917 (plain scaled dot-product branch)
  fairseq/modules/sinusoidal_positional_embedding.py (table + right-padded positions)
  espresso/criterions/label_smoothed_cross_entropy_v2.py:82-120 (uniform smoothing)
Pinned against the real reference by oracle/pin_against_reference.py (section "encdec").
"""
import math

import torch
import torch.nn.functional as F


def sinusoidal_table(n, d, padding_idx):
    half = d // 2
    f = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    e = torch.arange(n, dtype=torch.float)[:, None] * f[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if d % 2 == 1:
        e = torch.cat([e, torch.zeros(n, 1)], dim=1)
    if padding_idx is not None:
        e[padding_idx] = 0
    return e


def _drop(x, p, training):
    return F.dropout(x, p, training) if (training and p > 0) else x


def mha(sd, p, q_in, kv_in, H, key_pad, causal, attn_drop, training):
    """q_in [B, Tq, d], kv_in [B, Tk, d]; key_pad bool [B, Tk] or None."""
    B, Tq, d = q_in.shape
    Tk = kv_in.shape[1]
    hd = d // H
    q = F.linear(q_in, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * hd ** -0.5
    k = F.linear(kv_in, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(kv_in, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q = q.view(B, Tq, H, hd).transpose(1, 2)
    k = k.view(B, Tk, H, hd).transpose(1, 2)
    v = v.view(B, Tk, H, hd).transpose(1, 2)
    w = q @ k.transpose(-1, -2)
    if causal:
        w = w + torch.triu(torch.full((Tq, Tk), float("-inf"), device=w.device), 1)
    if key_pad is not None:
        w = w.masked_fill(key_pad[:, None, None, :], float("-inf"))
    w = _drop(torch.softmax(w.float(), dim=-1).to(w.dtype), attn_drop, training)
    ctx = (w @ v).transpose(1, 2).reshape(B, Tq, d)
    return F.linear(ctx, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def decoder_forward(sd, cfg, prev_output_tokens, enc_out, enc_pad, training=False, pre="decoder."):
    """prev_output_tokens [B, U]; enc_out [B, T', d]; enc_pad bool [B, T'] or None -> logits [B, U, V]."""
    pad = cfg["pad"]
    d, H = cfg["dec_embed_dim"], cfg["dec_heads"]
    dr, adr, acdr = cfg["dropout"], cfg["attention_dropout"], cfg["activation_dropout"]
    B, U = prev_output_tokens.shape
    x = math.sqrt(d) * F.embedding(prev_output_tokens, sd[pre + "embed_tokens.weight"], padding_idx=pad)
    nonpad = prev_output_tokens.ne(pad)
    positions = (torch.cumsum(nonpad.long(), dim=1) * nonpad.long() + pad)
    table = sinusoidal_table(U + pad + 2, d, pad).to(x)
    x = x + table[positions]
    if cfg.get("dec_layernorm_embedding", False):
        x = F.layer_norm(x, (d,), sd[pre + "layernorm_embedding.weight"], sd[pre + "layernorm_embedding.bias"])
    x = _drop(x, dr, training)
    tok_pad = prev_output_tokens.eq(pad)
    self_pad = tok_pad if bool(tok_pad.any()) else None
    for i in range(cfg["dec_layers"]):
        p = pre + "layers.%d." % i
        h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
        x = x + _drop(mha(sd, p + "self_attn.", h, h, H, self_pad, True, adr, training), dr, training)
        if enc_out is not None:  # enc_out None: decoder-only language model (fairseq transformer_lm layers)
            h = F.layer_norm(x, (d,), sd[p + "encoder_attn_layer_norm.weight"], sd[p + "encoder_attn_layer_norm.bias"])
            x = x + _drop(mha(sd, p + "encoder_attn.", h, enc_out, H, enc_pad, False, adr, training), dr, training)
        h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
        h = _drop(F.relu(F.linear(h, sd[p + "fc1.weight"], sd[p + "fc1.bias"])), acdr, training)
        x = x + _drop(F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"]), dr, training)
    x = F.layer_norm(x, (d,), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])
    wout = sd[pre + "embed_tokens.weight"] if cfg.get("share_decoder_input_output_embed", False) else sd[pre + "output_projection.weight"]
    return F.linear(x, wout)


def label_smoothed_ce(logits, target, eps, pad):
    """Uniform label smoothing, sum-reduced: returns (loss, nll_loss)."""
    lp = F.log_softmax(logits.float(), dim=-1).view(-1, logits.size(-1))
    t = target.reshape(-1)
    nll = -lp.gather(1, t[:, None]).squeeze(1)
    smooth = -lp.sum(-1)
    keep = t.ne(pad)
    nll, smooth = nll[keep].sum(), smooth[keep].sum()
    eps_i = eps / (lp.size(-1) - 1)
    return (1.0 - eps - eps_i) * nll + eps_i * smooth, nll
