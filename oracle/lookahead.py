"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the look-ahead word-LM fusion maths, never imported by the product.

Follows espresso/models/external_language_model.py:60-300 (`_LookAheadWordLanguageModelDecoder`, the per-hypothesis Python
version) and espresso/models/tensorized_lookahead_language_model.py:84-262 (the tensorised version the recipes use); both
implement Eqn. 15 of arXiv:1808.02608.  Pinned against the real TensorizedLookaheadLanguageModel by
oracle/pin_against_reference.py::pin_lookahead (tests/golden/lookahead_lm.npz).

The word LM itself is NOT restated here: the caller passes, per step, the word distribution the LM produced for every
hypothesis; this module owns the prefix tree, the state carried between steps and the subword log-probabilities."""
import numpy as np

ZERO = 1e-10


class Node:
    __slots__ = ("children", "word", "lo", "hi")

    def __init__(self):
        self.children, self.word, self.lo, self.hi = {}, -1, None, None


def build_tree(words, word_specials, subword_index, subword_unk, tokenizer=list):
    """words: list of strings by word id; word_specials: ids to skip; subword_index: str -> id.
    espresso/tools/lexical_prefix_tree.py:13-67."""
    root = Node()
    for w, text in enumerate(words):
        if w in word_specials:
            continue
        ids = [subword_index(s) for s in tokenizer(text)]
        if any(i == subword_unk for i in ids):
            continue
        node = root
        for i in ids:
            nxt = node.children.get(i)
            if nxt is None:
                nxt = node.children[i] = Node()
                nxt.lo, nxt.hi = w - 1, w
            else:
                nxt.lo, nxt.hi = min(nxt.lo, w - 1), max(nxt.hi, w)
            node = nxt
        node.word = w
    return root


class LookaheadState:
    def __init__(self, root, n):
        self.root, self.nodes, self.cum = root, [root] * n, None

    def reorder(self, order):
        """reorder_incremental_state (:264-278)."""
        self.nodes = [self.nodes[i] for i in order]
        self.cum = self.cum[np.asarray(order)]

    def lm_words(self, word_unk):
        """The word fed to the word LM at a non-first step: the word that ends at the current node, else <unk> (:126-130)."""
        return np.array([nd.word if nd is not None and nd.word >= 0 else word_unk for nd in self.nodes])


def step(state, prev_tokens, lm_probs, first, Vs, space, eos, pad, word_unk, word_eos, oov_penalty=1e-4, open_vocab=True):
    """prev_tokens [N]; lm_probs [N, Vw] = word LM softmax for the words of `state.lm_words` (or </s> at the first step).
    Updates state; returns log-probs [N, Vs] (float64 arithmetic on float32 cumulative sums, as the reference keeps them)."""
    N = len(prev_tokens)
    is_space = np.asarray(prev_tokens) == space
    new_cum = np.cumsum(lm_probs.astype(np.float32), axis=-1, dtype=np.float32)
    if first:
        state.cum = new_cum
        state.nodes = [state.root] * N
    else:
        state.cum = np.where(is_space[:, None], new_cum, state.cum)                      # :144
        for n in range(N):                                                                 # :150-164
            if is_space[n]:
                state.nodes[n] = state.root
            elif state.nodes[n] is not None and int(prev_tokens[n]) in state.nodes[n].children:
                state.nodes[n] = state.nodes[n].children[int(prev_tokens[n])]
            else:
                state.nodes[n] = None
    out = np.empty((N, Vs), dtype=np.float64)
    for n in range(N):
        cs, node = state.cum[n].astype(np.float64), state.nodes[n]
        if open_vocab:                                                                     # :173-197
            row = np.full(Vs, oov_penalty * (cs[word_unk] - cs[word_unk - 1]))
            if is_space[n] or prev_tokens[n] == eos:
                row[space] = ZERO
            if not is_space[n]:
                row[eos] = ZERO
            if node is None:
                row[:] = 1.0
        else:
            row = np.full(Vs, ZERO)
        sum_p = 1.0 if node is None or node is state.root else cs[node.hi] - cs[node.lo]   # :199-209
        if node is not None:
            for tok, child in node.children.items():                                       # :211-233
                row[tok] = ZERO if sum_p < ZERO else (cs[child.hi] - cs[child.lo]) / sum_p
        row[pad] = ZERO
        if node is not None and node.word >= 0:                                            # :236-255
            row[space] = ZERO if sum_p < ZERO else (cs[node.word] - cs[node.word - 1]) / sum_p
        out[n] = np.log(np.maximum(row, ZERO))
        if is_space[n]:                                                                    # :260-263
            out[n, eos] = np.log(lm_probs[n, word_eos])
    return out


# ---- multi-level (subword + word) LM: espresso/models/external_language_model.py:385-555 ---------------------------------
LOGZERO = -10.0


class MultiLevelState:
    def __init__(self, root, n):
        self.root, self.nodes, self.wlp, self.out, self.cumlp = root, [root] * n, None, None, np.zeros(n)

    def reorder(self, order):
        """:557-566"""
        o = np.asarray(order)
        self.nodes = [self.nodes[i] for i in order]
        self.wlp, self.out, self.cumlp = self.wlp[o], self.out[o], self.cumlp[o]

    def lm_words(self, word_unk):
        return np.array([nd.word if nd is not None and nd.word >= 0 else word_unk for nd in self.nodes])


def multilevel_step(state, prev_tokens, word_logprobs, sub_logprobs, first, space, eos, word_unk, word_eos, sub_weight=0.8,
                    oov_penalty=1.0, open_vocab=True):
    """word_logprobs [N, Vw]: word LM log-softmax for `state.lm_words` (</s> at the first step); sub_logprobs [N, Vs]: subword
    LM log-softmax for prev_tokens.  Returns the [N, Vs] row the search adds (already scaled by sub_weight)."""
    N = len(prev_tokens)
    is_space = np.asarray(prev_tokens) == space
    if first:
        state.wlp = word_logprobs.astype(np.float64).copy()
        state.cumlp = np.zeros(N)
        state.nodes = [state.root] * N
        is_child = np.zeros(N, dtype=bool)
    else:
        state.wlp = np.where(is_space[:, None], word_logprobs, state.wlp)                 # :425-436
        is_child = np.zeros(N, dtype=bool)
        for n in range(N):                                                                  # :437-455
            t = int(prev_tokens[n])
            if is_space[n]:
                state.nodes[n] = state.root
            elif state.nodes[n] is not None and t in state.nodes[n].children:
                state.nodes[n] = state.nodes[n].children[t]
                is_child[n] = True
            else:
                state.nodes[n] = None
        gathered = state.out[np.arange(N), np.asarray(prev_tokens)]
        if open_vocab:                                                                      # :456-470
            state.cumlp = np.where(is_space, 0.0, state.cumlp + gathered)
        else:
            state.cumlp = np.where(is_child, state.cumlp + gathered, 0.0)
    out = sub_logprobs.astype(np.float64) * sub_weight                                      # :472-481
    if not open_vocab and not first:
        out[~is_space & ~is_child] = LOGZERO                                                # :483-485
    w = state.lm_words(word_unk)                                                            # :497-509
    wl = state.wlp[np.arange(N), w]
    wl = wl + np.where(w != word_unk, -state.cumlp, np.log(oov_penalty))                    # :510-516
    out[:, space] = wl
    out[is_space | (np.asarray(prev_tokens) == eos), space] = LOGZERO                       # :521-526
    out[~is_space, eos] = LOGZERO
    out[is_space, eos] += state.wlp[is_space, word_eos]                                     # :529-532
    state.out = out
    return out
