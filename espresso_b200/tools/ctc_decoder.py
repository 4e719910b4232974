"""Greedy CTC decoder used for validation WER (espresso/tools/ctc_decoder.py:21-188 without the optional KenLM path):
argmax over the vocabulary per frame, collapse repeats (unique_consecutive), drop blanks.
decode(models, sample) -> (tokens int64 [B, U] padded, scores None, alignments None), like the reference's
`decode` (:80-98).  The argmax runs in esp_argmax_rows; the collapse is a tiny host loop over B sequences."""
import torch

from .. import ops as _ops


class CTCDecoder:
    def __init__(self, dictionary, blank_idx=None, pad_idx=None, **unused):
        self.pad = dictionary.pad() if pad_idx is None else pad_idx
        self.blank = dictionary.index("<s>") if blank_idx is None else blank_idx

    @torch.no_grad()
    def decode(self, models, sample, **unused):
        model = models[0]
        model.eval()
        net = model(**sample["net_input"])
        out = net["b200_out"]  # [B, T', ld]
        V = net["encoder_out"][0].size(-1)
        B, T, ld = out.shape
        lens = net["src_lengths"][0].tolist()
        am = _ops.argmax_rows(out.reshape(B * T, ld), V).view(B, T).cpu()
        hyps = []
        for b in range(B):
            seq = am[b, : lens[b]]
            if seq.numel():
                keep = torch.ones_like(seq, dtype=torch.bool)
                keep[1:] = seq[1:] != seq[:-1]          # unique_consecutive
                seq = seq[keep]
                seq = seq[seq != self.blank]
            hyps.append(seq.long())
        U = max([h.numel() for h in hyps] + [1])
        tokens = torch.full((B, U), self.pad, dtype=torch.long)
        for b, h in enumerate(hyps):
            tokens[b, : h.numel()] = h
        return tokens, None, None

    def generate(self, models, sample, **unused):
        tokens, _, _ = self.decode(models, sample)
        return [[{"tokens": t[t != self.pad], "score": 0.0, "attention": None, "alignment": None, "positional_scores": None}]
                for t in tokens]
