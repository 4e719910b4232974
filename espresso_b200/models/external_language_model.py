"""Multi-level (subword + word) LM shallow fusion (espresso/models/external_language_model.py:306-587; Hori et al.,
"Multi-level language modeling and decoding for open vocabulary end-to-end speech recognition", ASRU 2017).

The subword LM scores every search step; at a word boundary the word LM's log-probability of the completed word replaces what
the subword LM accumulated inside that word, out-of-lexicon words fall back to the word LM's <unk> plus a penalty.  Same
B200 structure as the look-ahead LM (tensorized_lookahead_language_model.py): the word LM's rows live on the device, the
prefix tree is CSR, and the per-hypothesis Python loops of the reference (:437-455, :497-509, node lists re-built on every
reorder :562-566) are one launch of csrc/lookahead.cu::multilevel_step_kernel."""
import math

import torch
import torch.nn as nn

from .. import ops as _ops
from ..tools.tensorized_prefix_tree import TensorizedPrefixTree


class MultiLevelLanguageModel(nn.Module):
    def __init__(self, wordlm, subwordlm, subwordlm_weight=0.8, oov_penalty=1.0, open_vocab=True):
        super().__init__()
        dec = getattr(wordlm, "decoder", None)
        if dec is None or not all(hasattr(dec, a) for a in ("embed_tokens", "step", "output_layer", "layers", "hidden_size")):
            raise TypeError("the word LM must be an lstm_lm_espresso model (masked_copy_cached_state, external_language_model.py:349-352)")
        if not hasattr(subwordlm, "decode_step"):
            raise TypeError("the subword LM must speak the generator protocol (init_incremental_state / decode_step)")
        self.wordlm, self.subwordlm = wordlm, subwordlm
        self.subwordlm_weight, self.log_oov_penalty, self.open_vocab, self.logzero = float(subwordlm_weight), math.log(oov_penalty), bool(open_vocab), -10.0
        wd = dec.dictionary
        sd = getattr(subwordlm, "dictionary", None) or subwordlm.decoder.dictionary
        self.word_eos, self.word_unk, self.n_words = wd.eos(), wd.unk(), len(wd)
        self.space, self.eos, self.vocab = sd.space(), sd.eos(), len(sd)
        if self.space < 0:
            raise ValueError("the subword dictionary has no space symbol")
        self.tree = TensorizedPrefixTree.build(wd, sd)

    def finalize_(self, device, dtype=torch.bfloat16):
        self.wordlm.finalize_(device, dtype)
        self.subwordlm.finalize_(device, dtype)
        self.tree.to(device)
        return self

    def max_decoder_positions(self):
        return int(1e5)  # external_language_model.py:573-574

    def forward_encoder(self, net_input):
        return None

    def init_incremental_state(self, encoder_out, bsz, beam):
        N = bsz * beam
        dec = self.wordlm.decoder
        w = dec.embed_tokens.weight
        dev = w.device
        ld = (self.vocab + 7) // 8 * 8
        i32 = lambda fill: torch.full((N,), fill, dtype=torch.int32, device=dev)  # noqa: E731
        f32 = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
        return {
            "sub": self.subwordlm.init_incremental_state(None, bsz, beam),
            "h": [w.new_zeros(N, dec.hidden_size) for _ in dec.layers], "c": [w.new_zeros(N, dec.hidden_size) for _ in dec.layers],
            "nodes": i32(TensorizedPrefixTree.root_id), "nodes_tmp": i32(0), "words": i32(self.word_eos),
            "wlp": f32(N, self.n_words), "wlp_alt": f32(N, self.n_words), "eos_lp": f32(N),
            "out": f32(N, ld), "out_alt": f32(N, ld), "cumlp": f32(N), "cumlp_alt": f32(N),
        }

    @torch.no_grad()
    def decode_step(self, step, tokens, state, new_order):
        dec = self.wordlm.decoder
        tree = self.tree.to(tokens.device)
        first = step == 0
        prev = tokens[:, step]
        sub_vals, sub_is_logits = self.subwordlm.decode_step(step, tokens, state["sub"], new_order)
        if first:
            state["words"].fill_(self.word_eos)
            h_old, c_old, nodes_in = state["h"], state["c"], state["nodes"]
        else:
            _ops.lookahead_words(state["nodes"], new_order, tree["node_word"], self.word_unk, state["nodes_tmp"], state["words"])
            nodes_in = state["nodes_tmp"]
            idx = None if new_order is None else new_order.long()
            h_old = state["h"] if idx is None else [h.index_select(0, idx) for h in state["h"]]
            c_old = state["c"] if idx is None else [c.index_select(0, idx) for c in state["c"]]
        y, h_new, c_new, _ = dec.step(dec.embed_tokens(state["words"].long()), h_old, c_old, None)
        logits = dec.output_layer(y)
        if first:
            state["h"], state["c"] = h_new, c_new
        else:
            fresh = (prev == self.space)[:, None]
            state["h"] = [torch.where(fresh, a, b) for a, b in zip(h_new, h_old)]
            state["c"] = [torch.where(fresh, a, b) for a, b in zip(c_new, c_old)]
        # word-level log-probabilities: refreshed after a <space>, inherited from the parent hypothesis otherwise (:425-436)
        _ops.wordlm_cumsum(logits, self.n_words, prev, tokens.stride(0), self.space, first, state["wlp"], new_order, state["wlp_alt"],
                           state["eos_lp"], self.word_eos, log_mode=True)
        state["wlp"], state["wlp_alt"] = state["wlp_alt"], state["wlp"]
        _ops.multilevel_step(prev, tokens.stride(0), first, nodes_in, state["nodes"], new_order, state["wlp"], self.n_words, sub_vals,
                             sub_is_logits, self.subwordlm_weight, state["out"], state["cumlp"], state["cumlp_alt"], tree, self.space,
                             self.eos, self.word_unk, self.word_eos, self.log_oov_penalty, self.open_vocab, self.logzero,
                             state["out_alt"], self.vocab)
        state["out"], state["out_alt"] = state["out_alt"], state["out"]
        state["cumlp"], state["cumlp_alt"] = state["cumlp_alt"], state["cumlp"]
        return state["out"], False


def wrap_language_models(lms, subword_dict, subwordlm_weight=0.8, oov_penalty=1e-4, open_vocab=True):
    """The LM set-up of espresso/speech_recognize.py:132-160: `lms` are the models loaded from --lm-path (one or two).  A
    word LM (`is_wordlm`) preceded by a subword LM becomes a MultiLevelLanguageModel; a lone word LM becomes a
    TensorizedLookaheadLanguageModel; a subword LM is used as it is.  Returns the single `lm_model=` for the generator."""
    from .tensorized_lookahead_language_model import TensorizedLookaheadLanguageModel

    lms = [m for m in lms if m is not None]
    if not 1 <= len(lms) <= 2:
        raise ValueError("expected one LM, or a subword LM followed by a word LM")
    for i, m in enumerate(lms):
        if getattr(m, "is_wordlm", False):
            if i > 0:
                return MultiLevelLanguageModel(m, lms[i - 1], subwordlm_weight=subwordlm_weight, oov_penalty=oov_penalty,
                                               open_vocab=open_vocab)
            if len(lms) > 1:
                raise ValueError("the subword LM must come before the word LM")
            return TensorizedLookaheadLanguageModel(m, subword_dict, oov_penalty=oov_penalty, open_vocab=open_vocab)
    if len(lms) != 1:
        raise ValueError("two LMs were given but neither is a word LM (is_wordlm)")
    return lms[0]
