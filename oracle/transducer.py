"""Oracle: transducer prediction network + joint + RNN-T loss, functional PyTorch on the reference's parameter names.
TEST INFRASTRUCTURE ONLY.  Restates espresso/models/speech_lstm.py:766-919 (SpeechLSTMDecoder.extract_features used as
RNN-T predictor: embedding -> LSTMCell stack, zero initial state, no attention),
espresso/models/transformer/speech_transformer_transducer_base.py:255-299 (joint, weight-normalised fc_out) and
espresso/criterions/transducer_loss.py:73-154.  Pinned by oracle/pin_against_reference.py (section "transducer")."""
import torch
import torch.nn.functional as F


def predictor(sd, prev_output_tokens, n_layers, pad, pre="decoder."):
    x = F.embedding(prev_output_tokens, sd[pre + "embed_tokens.weight"], padding_idx=pad)
    B, U1, _ = x.shape
    for i in range(n_layers):
        p = pre + "layers.%d." % i
        Hd = sd[p + "weight_hh"].shape[1]
        h, c = x.new_zeros(B, Hd), x.new_zeros(B, Hd)
        outs = []
        for u in range(U1):
            gates = F.linear(x[:, u], sd[p + "weight_ih"], sd[p + "bias_ih"]) + F.linear(h, sd[p + "weight_hh"], sd[p + "bias_hh"])
            i_, f_, g_, o_ = gates.chunk(4, dim=1)
            c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            outs.append(h)
        x = torch.stack(outs, dim=1)
    return x


def joint_logits(sd, enc, dec):
    J = sd["proj_encoder.weight"].shape[0]
    pe = F.layer_norm(F.linear(enc, sd["proj_encoder.weight"], sd["proj_encoder.bias"]), (J,), sd["laynorm_proj_encoder.weight"],
                      sd["laynorm_proj_encoder.bias"])
    pd = F.layer_norm(F.linear(dec, sd["proj_decoder.weight"], sd["proj_decoder.bias"]), (J,), sd["laynorm_proj_decoder.weight"],
                      sd["laynorm_proj_decoder.bias"])
    f = F.relu(pe[:, :, None, :] + pd[:, None, :, :])
    v, g = sd["fc_out.weight_v"], sd["fc_out.weight_g"]
    W = g * v / v.norm(dim=1, keepdim=True)
    return F.linear(f, W, sd["fc_out.bias"])


def transducer_loss(logits, enc_lens, target, pad, eos, blank):
    import torchaudio

    u_lens = ((target != pad) & (target != eos)).sum(-1).int()
    return torchaudio.functional.rnnt_loss(logits.float(), target[:, :-1].int().contiguous(), enc_lens.int(), u_lens, blank=blank,
                                           clamp=-1.0, reduction="sum")


def lstm_lm_step(lm_sd, prev, hs, cs, pad):
    """One step of the attention-free LSTM LM (espresso/models/lstm_lm.py -> SpeechLSTMDecoder without encoder):
    returns (log-probs [B, V_lm], new hs, new cs)."""
    x = F.embedding(prev, lm_sd["decoder.embed_tokens.weight"], padding_idx=pad)
    nh, nc = [], []
    for i in range(len(hs)):
        p = "decoder.layers.%d." % i
        gates = F.linear(x, lm_sd[p + "weight_ih"], lm_sd[p + "bias_ih"]) + F.linear(hs[i], lm_sd[p + "weight_hh"], lm_sd[p + "bias_hh"])
        i_, f_, g_, o_ = gates.chunk(4, dim=1)
        c = torch.sigmoid(f_) * cs[i] + torch.sigmoid(i_) * torch.tanh(g_)
        h = torch.sigmoid(o_) * torch.tanh(c)
        nh.append(h)
        nc.append(c)
        x = h
    if "decoder.additional_fc.weight" in lm_sd:
        x = F.linear(x, lm_sd["decoder.additional_fc.weight"], lm_sd["decoder.additional_fc.bias"])
    if "decoder.fc_out.weight" in lm_sd:
        logits = F.linear(x, lm_sd["decoder.fc_out.weight"], lm_sd["decoder.fc_out.bias"])
    else:
        logits = F.linear(x, lm_sd["decoder.embed_tokens.weight"])
    return torch.log_softmax(logits, dim=-1), nh, nc


def greedy_decode(sd, enc, enc_lens, n_layers, pad, blank, bos, eos, max_num_expansions_per_step=2, temperature=1.0,
                  model_predicts_eos=False, max_len=0, lm_sd=None, lm_weight=1.0):
    """Restates espresso/tools/transducer_greedy_decoder.py:91-251 (no LM): frame-synchronous greedy search, at most
    `max_num_expansions_per_step` non-blank tokens per encoder frame, predictor state rolled back on blank.
    enc [B, T, d] (eval-mode encoder output), enc_lens [B].
    Returns (tokens int64 [B, T*(E+1)], scores [B], margins [B, T, E+1] = top-1 minus top-2 log-prob of every decision,
    +inf where no decision was taken)."""
    B, T, _ = enc.shape
    T = min(int(enc_lens.max()), max_len) if max_len > 0 else int(enc_lens.max())
    E = max_num_expansions_per_step
    tokens = torch.full((B, T, E + 1), blank, dtype=torch.long)
    scores = torch.zeros(B, T, E + 1)
    margins = torch.full((B, T, E + 1), float("inf"))
    prev = torch.full((B,), bos, dtype=torch.long)
    hid = [sd["decoder.layers.%d.weight_hh" % i].shape[1] for i in range(n_layers)]
    hs = [torch.zeros(B, h) for h in hid]
    cs = [torch.zeros(B, h) for h in hid]
    if lm_sd is not None:  # LM shallow fusion (transducer_greedy_decoder.py:165-201), LM vocabulary = ASR vocabulary
        n_lm = len([k for k in lm_sd if k.endswith("weight_hh")])
        lhs = [torch.zeros(B, lm_sd["decoder.layers.%d.weight_hh" % i].shape[1]) for i in range(n_lm)]
        lcs = [torch.zeros_like(h) for h in lhs]
        nonblank = torch.ones(sd["fc_out.bias"].numel(), dtype=torch.bool)
        nonblank[blank] = False
    for t in range(T):
        blank_mask = t >= enc_lens
        k = 0
        while not bool(blank_mask.all()) and k < E + 1:
            # one predictor step from the cached state
            x = F.embedding(prev, sd["decoder.embed_tokens.weight"], padding_idx=pad)
            nh, nc = [], []
            for i in range(n_layers):
                p = "decoder.layers.%d." % i
                gates = F.linear(x, sd[p + "weight_ih"], sd[p + "bias_ih"]) + F.linear(hs[i], sd[p + "weight_hh"], sd[p + "bias_hh"])
                i_, f_, g_, o_ = gates.chunk(4, dim=1)
                c = torch.sigmoid(f_) * cs[i] + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                nh.append(h)
                nc.append(c)
                x = h
            logits = joint_logits(sd, enc[:, t:t + 1], x[:, None, :])[:, 0, 0]  # [B, V]
            lp = torch.log_softmax(logits / temperature, dim=-1)
            if lm_sd is not None:
                lm_lp, lnh, lnc = lstm_lm_step(lm_sd, prev, lhs, lcs, pad)
                lp_nb = lp[:, nonblank]
                fused = lp_nb + lm_weight * lm_lp[:, nonblank]
                # the non-blank probability mass stays what the transducer assigned
                fused = fused + (lp_nb.exp().sum(1).log() - fused.exp().sum(1).log())[:, None]
                lp = lp.clone()
                lp[:, nonblank] = fused
            if model_predicts_eos:
                lp[:, blank] = torch.logaddexp(lp[:, blank], lp[:, eos])
                lp[:, eos] = float("-inf")
            if k < E:
                top2 = lp.topk(2, dim=-1)
                sc, tok = top2.values[:, 0], top2.indices[:, 0]
                margins[~blank_mask, t, k] = (top2.values[:, 0] - top2.values[:, 1])[~blank_mask]
                scores[:, t, k] = torch.where(blank_mask, torch.zeros_like(sc), sc)
                blank_mask = blank_mask | (tok == blank)
                tokens[:, t, k] = torch.where(blank_mask, torch.full_like(tok, blank), tok)
                prev = torch.where(blank_mask, prev, tok)
            else:
                scores[:, t, k] = torch.where(blank_mask, scores[:, t, k], lp[:, blank])
                blank_mask = torch.ones_like(blank_mask)
            # rows that emitted blank (or are done with this frame) keep the old predictor state
            keep_old = blank_mask[:, None]
            hs = [torch.where(keep_old, o, n) for o, n in zip(hs, nh)]
            cs = [torch.where(keep_old, o, n) for o, n in zip(cs, nc)]
            if lm_sd is not None:
                lhs = [torch.where(keep_old, o, n) for o, n in zip(lhs, lnh)]
                lcs = [torch.where(keep_old, o, n) for o, n in zip(lcs, lnc)]
            k += 1
    return tokens.view(B, -1), scores.view(B, -1).sum(-1), margins


def search_callbacks(sd, enc_b, n_layers, pad, temperature=1.0, lm_sd=None):
    """fp32 model callbacks for espresso_b200.tools.transducer_beam_search_decoder.AdaptiveExpansionSearch (one utterance,
    enc_b [T, d]): the prediction network / joint / LM of this module behind the search's callback protocol, so the
    host search can be checked against the reference decoder without any bf16 noise."""
    hid = [sd["decoder.layers.%d.weight_hh" % i].shape[1] for i in range(n_layers)]
    assert len(set(hid)) == 1

    def cell_stack(prefix, weights, x, hs, cs, n):
        nh, nc = [], []
        for i in range(n):
            p = prefix + "layers.%d." % i
            gates = F.linear(x, weights[p + "weight_ih"], weights[p + "bias_ih"]) + F.linear(hs[:, i], weights[p + "weight_hh"], weights[p + "bias_hh"])
            i_, f_, g_, o_ = gates.chunk(4, dim=1)
            c = torch.sigmoid(f_) * cs[:, i] + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            nh.append(h)
            nc.append(c)
            x = h
        return x, torch.stack(nh, dim=1), torch.stack(nc, dim=1)

    def pred_step(prev, hs, cs):
        x = F.embedding(prev, sd["decoder.embed_tokens.weight"], padding_idx=pad)
        return cell_stack("decoder.", sd, x, hs, cs, n_layers)

    def joint_lprobs(t, out):
        n = out.size(0)
        logits = joint_logits(sd, enc_b[None, t:t + 1].expand(n, -1, -1), out[:, None, :])[:, 0, 0]
        return torch.log_softmax(logits / temperature, dim=-1)

    cb = dict(pred_step=pred_step, joint_lprobs=joint_lprobs,
              init_state=lambda n: (torch.zeros(n, n_layers, hid[0]), torch.zeros(n, n_layers, hid[0])))
    if lm_sd is not None:
        n_lm = len([k for k in lm_sd if k.endswith("weight_hh")])
        lh = lm_sd["decoder.layers.0.weight_hh"].shape[1]

        def lm_step(prev, lhs, lcs):
            x = F.embedding(prev, lm_sd["decoder.embed_tokens.weight"], padding_idx=pad)
            return cell_stack("decoder.", lm_sd, x, lhs, lcs, n_lm)

        def lm_lprobs(feat):
            x = feat
            if "decoder.additional_fc.weight" in lm_sd:
                x = F.linear(x, lm_sd["decoder.additional_fc.weight"], lm_sd["decoder.additional_fc.bias"])
            w = lm_sd["decoder.fc_out.weight"] if "decoder.fc_out.weight" in lm_sd else lm_sd["decoder.embed_tokens.weight"]
            b = lm_sd.get("decoder.fc_out.bias") if "decoder.fc_out.weight" in lm_sd else None
            return torch.log_softmax(F.linear(x, w, b), dim=-1)

        cb.update(lm_step=lm_step, lm_lprobs=lm_lprobs, lm_init_state=lambda n: (torch.zeros(n, n_lm, lh), torch.zeros(n, n_lm, lh)))
    return cb
