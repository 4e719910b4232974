"""Look-ahead word-LM shallow fusion (espresso/models/tensorized_lookahead_language_model.py:18-304; Hori et al.,
arXiv:1808.02608): a WORD language model scores the subword hypotheses of the beam search through a lexical prefix tree.

B200 design: the per-step work of the reference (~40 gather / scatter / where ops on dense tables, a clone of the whole LM
state and an index_select of the [N, |words|] cumulative distributions) is three launches of csrc/lookahead.cu around the
word LM's own step -- see the kernel file for the data layout.  The object speaks the generator protocol of
espresso_b200.sequence_generator (`init_incremental_state` / `decode_step`), returns LOG-PROBABILITIES
(`is_logits=False`, like the reference's RawOutExternalLanguageModelBase), and is passed as `lm_model=`."""
import torch
import torch.nn as nn

from .. import ops as _ops
from ..tools.tensorized_prefix_tree import TensorizedPrefixTree


class TensorizedLookaheadLanguageModel(nn.Module):
    def __init__(self, word_lm, subword_dict, oov_penalty=1e-4, open_vocab=True):
        super().__init__()
        dec = getattr(word_lm, "decoder", None)
        if dec is None or not all(hasattr(dec, a) for a in ("embed_tokens", "step", "output_layer", "layers", "hidden_size")):
            raise TypeError("the word LM must be an lstm_lm_espresso model (the reference requires masked_copy_cached_state, "
                            "which only its LSTM decoders implement: tensorized_lookahead_language_model.py:57-60)")
        self.word_lm = word_lm
        self.oov_penalty, self.open_vocab, self.zero = float(oov_penalty), bool(open_vocab), 1e-10
        wd = dec.dictionary
        self.word_eos, self.word_unk, self.n_words = wd.eos(), wd.unk(), len(wd)
        self.space, self.pad, self.eos, self.vocab = subword_dict.space(), subword_dict.pad(), subword_dict.eos(), len(subword_dict)
        if self.space < 0:
            raise ValueError("the subword dictionary has no space symbol")
        self.tree = TensorizedPrefixTree.build(wd, subword_dict)
        assert self.tree.max_out_degree() <= self.vocab

    def finalize_(self, device, dtype=torch.bfloat16):
        self.word_lm.finalize_(device, dtype)
        self.tree.to(device)
        return self

    def max_decoder_positions(self):
        return int(1e5)  # tensorized_lookahead_language_model.py:288-289

    def forward_encoder(self, net_input):
        return None

    def init_incremental_state(self, encoder_out, bsz, beam):
        N = bsz * beam
        dec = self.word_lm.decoder
        w = dec.embed_tokens.weight
        dev = w.device
        ld = (self.vocab + 7) // 8 * 8
        i32 = lambda fill: torch.full((N,), fill, dtype=torch.int32, device=dev)  # noqa: E731
        f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)  # noqa: E731
        return {
            "h": [w.new_zeros(N, dec.hidden_size) for _ in dec.layers], "c": [w.new_zeros(N, dec.hidden_size) for _ in dec.layers],
            "nodes": i32(TensorizedPrefixTree.root_id), "nodes_tmp": i32(0), "words": i32(self.word_eos),
            "cum": f32(N, self.n_words), "cum_alt": f32(N, self.n_words), "eos_lp": f32(N), "out": f32(N, ld),
        }

    @torch.no_grad()
    def decode_step(self, step, tokens, state, new_order):
        """tokens int32 [N, L]: column `step` is the subword every hypothesis has just emitted (</s> at step 0)."""
        dec = self.word_lm.decoder
        tree = self.tree.to(tokens.device)
        first = step == 0
        prev = tokens[:, step]
        if first:
            state["words"].fill_(self.word_eos)      # the word LM starts from </s> (:109-113); every node is the root
            h_old, c_old, nodes_in = state["h"], state["c"], state["nodes"]
        else:
            # beam reordering of the node ids + the word each hypothesis has just completed (:126-130), one launch
            _ops.lookahead_words(state["nodes"], new_order, tree["node_word"], self.word_unk, state["nodes_tmp"], state["words"])
            nodes_in = state["nodes_tmp"]
            idx = None if new_order is None else new_order.long()
            h_old = state["h"] if idx is None else [h.index_select(0, idx) for h in state["h"]]
            c_old = state["c"] if idx is None else [c.index_select(0, idx) for c in state["c"]]
        x = dec.embed_tokens(state["words"].long())
        y, h_new, c_new, _ = dec.step(x, h_old, c_old, None)
        logits = dec.output_layer(y)
        if first:
            state["h"], state["c"] = h_new, c_new
        else:
            # the LM state only advances across a word boundary (masked_copy_cached_state, :139-143)
            fresh = (prev == self.space)[:, None]
            state["h"] = [torch.where(fresh, a, b) for a, b in zip(h_new, h_old)]
            state["c"] = [torch.where(fresh, a, b) for a, b in zip(c_new, c_old)]
        _ops.wordlm_cumsum(logits, self.n_words, prev, tokens.stride(0), self.space, first, state["cum"], new_order,
                           state["cum_alt"], state["eos_lp"], self.word_eos)
        state["cum"], state["cum_alt"] = state["cum_alt"], state["cum"]
        # tree transition + subword log-probabilities (each CTA touches only its own node id: in place is fine)
        _ops.lookahead_step(prev, tokens.stride(0), first, nodes_in, state["nodes"], state["cum"], self.n_words, state["eos_lp"],
                            tree, self.space, self.eos, self.pad, self.word_unk, self.oov_penalty, self.open_vocab, self.zero,
                            state["out"], self.vocab)
        return state["out"], False
