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
