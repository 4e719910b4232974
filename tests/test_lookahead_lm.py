"""Look-ahead word-LM fusion on the host (no GPU): the oracle restatement against outputs recorded from the REAL
TensorizedLookaheadLanguageModel (tests/golden/lookahead_lm.npz, oracle/pin_against_reference.py::pin_lookahead), and the
product's CSR prefix tree against the oracle's pointer tree."""
import os

import numpy as np
import pytest


def _dicts(g):
    from espresso_b200.data.asr_dictionary import AsrDictionary

    sub, wrd = AsrDictionary(), AsrDictionary()
    for c in list(g["chars"]) + ["<space>"]:
        sub.add_symbol(str(c))
    sub.space_index = sub.indices.get(sub.space_word, -1)
    for w in g["words"]:
        wrd.add_symbol(str(w))
    return sub, wrd


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "lookahead_lm.npz"))


@pytest.mark.parametrize("variant,open_vocab,pen", [("open", True, 1e-4), ("closed", False, 1e-4), ("open_pen", True, 0.3)])
def test_oracle_reproduces_reference_outputs(g, variant, open_vocab, pen):
    from oracle import lookahead as OL

    sub, wrd = _dicts(g)
    root = OL.build_tree([wrd[i] for i in range(len(wrd))], {wrd.pad(), wrd.eos(), wrd.unk()}, sub.index, sub.unk())
    prev, orders, ref, lm = (g[variant + k] for k in (".prev_tokens", ".new_orders", ".out", ".lm_probs"))
    state = OL.LookaheadState(root, prev.shape[1])
    for t in range(prev.shape[0]):
        o = OL.step(state, prev[t], lm[t], t == 0, len(sub), sub.space(), sub.eos(), sub.pad(), wrd.unk(), wrd.eos(), pen, open_vocab)
        big = ref[t] > -15
        assert np.abs(o - ref[t]).max() < 1.0 and (not big.any() or np.abs(o - ref[t])[big].max() < 2e-4), (variant, t)
        if t < len(orders):
            state.reorder(orders[t])
    # the recorded run exercised every case of Eqn. 15: in-lexicon continuation, word end, OOV run, sentence end
    assert (ref[:, :, sub.space()] > -5).any() and (ref[:, :, sub.eos()] > -20).any()
    if open_vocab:   # a hypothesis that left the lexicon runs free: log-prob 0 for every subword but <pad>
        free = np.delete(ref, sub.pad(), axis=-1)
        assert (np.abs(free).max(-1) == 0).any()


def test_csr_tree_matches_pointer_tree(g):
    from espresso_b200.tools.tensorized_prefix_tree import TensorizedPrefixTree
    from oracle import lookahead as OL

    sub, wrd = _dicts(g)
    tree = TensorizedPrefixTree.build(wrd, sub)
    root = OL.build_tree([wrd[i] for i in range(len(wrd))], {wrd.pad(), wrd.eos(), wrd.unk()}, sub.index, sub.unk())
    assert tree.child_off[0] == tree.child_off[1] == 0                      # node 0 ("outside the lexicon") has no edges
    n_nodes = [0]

    def walk(node, k):
        n_nodes[0] += 1
        e0, e1 = tree.child_off[k], tree.child_off[k + 1]
        toks = tree.child_tok[e0:e1].tolist()
        assert toks == sorted(node.children) and len(set(toks)) == len(toks)
        if node is not root:
            assert (int(tree.node_word[k]), int(tree.node_lo[k]), int(tree.node_hi[k])) == (node.word, node.lo, node.hi)
        for tok, child in node.children.items():
            assert tree.step(k, tok) == tree.child_node[e0 + toks.index(tok)]
            walk(child, tree.step(k, tok))

    walk(root, tree.root_id)
    assert n_nodes[0] == tree.num_nodes - 1                                   # every node reachable, plus node 0
    assert tree.max_out_degree() == len(root.children)
    assert tree.step(tree.root_id, sub.unk()) == tree.none_id and tree.step(tree.none_id, 5) == tree.none_id
    # "quiz" has characters outside the subword set: not in the tree; every other word is found by spelling it
    for w in range(len(wrd)):
        if w in (wrd.pad(), wrd.eos(), wrd.unk()):
            continue
        k = tree.root_id
        for ch in wrd[w]:
            k = tree.step(k, sub.index(ch)) if k else 0
        assert (k != 0 and tree.node_word[k] == w) or wrd[w] == "quiz"
    assert tree.node_lo.dtype == np.int32 and (tree.node_hi[2:] > tree.node_lo[2:]).all()


def test_tree_rejects_dictionaries_without_a_special_at_zero(g):
    from espresso_b200.data.asr_dictionary import AsrDictionary
    from espresso_b200.tools.tensorized_prefix_tree import TensorizedPrefixTree

    sub, _ = _dicts(g)
    wrd = AsrDictionary(enable_bos=True)
    wrd.add_symbol("a")
    with pytest.raises(ValueError):
        TensorizedPrefixTree.build(wrd, sub)


@pytest.mark.parametrize("variant,open_vocab,pen", [("open", True, 1e-4), ("closed", False, 1e-4), ("open_pen", True, 0.3)])
def test_wrapper_host_logic_vs_reference_fixture(g, variant, open_vocab, pen, monkeypatch):
    """TensorizedLookaheadLanguageModel.decode_step (state ping-pong, masked LM state update, reordering) with the three
    kernels replaced by their oracle statements: the same sequence the GPU test runs through the CUDA path."""
    import torch

    from espresso_b200 import ops
    from espresso_b200.models import LSTMLanguageModelEspresso, LSTMLanguageModelEspressoConfig, TensorizedLookaheadLanguageModel
    from oracle import ops_ref

    for name in ("lookahead_words", "wordlm_cumsum", "lookahead_step"):
        monkeypatch.setattr(ops, name, getattr(ops_ref, name))
    sub, wrd = _dicts(g)

    class _Task:
        target_dictionary = wrd

    e, h, o, nl = (int(v) for v in g["lm_cfg"])
    lm = LSTMLanguageModelEspresso.build_model(LSTMLanguageModelEspressoConfig(
        dropout=0.0, decoder_embed_dim=e, decoder_hidden_size=h, decoder_layers=nl, decoder_out_embed_dim=o), _Task())
    lm.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    m = TensorizedLookaheadLanguageModel(lm.eval(), sub, oov_penalty=pen, open_vocab=open_vocab)
    prev, orders, ref = (g[variant + k] for k in (".prev_tokens", ".new_orders", ".out"))
    S, N = prev.shape
    tokens = torch.from_numpy(prev.T.copy()).to(torch.int32)
    state = m.init_incremental_state(None, N, 1)
    for t in range(S):
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32)
        out, is_logits = m.decode_step(t, tokens, state, order)
        got = out[:, : len(sub)].numpy()
        big = ref[t] > -15
        assert not is_logits and np.abs(got - ref[t]).max() < 1.0
        assert not big.any() or np.abs(got - ref[t])[big].max() < 2e-4, (variant, t)


# ---- multi-level (subword + word) LM -----------------------------------------------------------------------------------
ML_VARIANTS = ["open", "open_pen", "open_w1"]   # closed vocabulary: the reference raises TypeError (external_language_model.py:484)


@pytest.fixture(scope="module")
def gm(golden_dir):
    return np.load(os.path.join(golden_dir, "multilevel_lm.npz"))


def _lstm_lm(gm, prefix, cfg_key, d):
    import torch

    from espresso_b200.models import LSTMLanguageModelEspresso, LSTMLanguageModelEspressoConfig

    class _Task:
        target_dictionary = d

    e, h, o, nl = (int(v) for v in gm[cfg_key])
    lm = LSTMLanguageModelEspresso.build_model(LSTMLanguageModelEspressoConfig(
        dropout=0.0, decoder_embed_dim=e, decoder_hidden_size=h, decoder_layers=nl, decoder_out_embed_dim=o), _Task())
    lm.load_state_dict({k[len(prefix):]: torch.from_numpy(gm[k]) for k in gm.files if k.startswith(prefix)}, strict=True)
    return lm.eval()


@pytest.mark.parametrize("variant", ML_VARIANTS)
def test_multilevel_oracle_reproduces_reference_outputs(gm, variant):
    from oracle import lookahead as OL

    sub, wrd = _dicts(gm)
    root = OL.build_tree([wrd[i] for i in range(len(wrd))], {wrd.pad(), wrd.eos(), wrd.unk()}, sub.index, sub.unk())
    prev, orders, ref, wl, sl = (gm[variant + k] for k in (".prev_tokens", ".new_orders", ".out", ".word_logprobs", ".sub_logprobs"))
    open_vocab, pen, weight = gm[variant + ".params"]
    state = OL.MultiLevelState(root, prev.shape[1])
    for t in range(prev.shape[0]):
        o = OL.multilevel_step(state, prev[t], wl[t], sl[t], t == 0, sub.space(), sub.eos(), wrd.unk(), wrd.eos(), float(weight), float(pen),
                               bool(open_vocab))
        assert np.abs(o - ref[t]).max() < 1e-4, (variant, t)
        if t < len(orders):
            state.reorder(orders[t])


@pytest.mark.parametrize("variant", ML_VARIANTS)
def test_multilevel_wrapper_host_logic_vs_reference_fixture(gm, variant, monkeypatch):
    import torch

    from espresso_b200 import ops
    from espresso_b200.models import MultiLevelLanguageModel
    from oracle import ops_ref

    for name in ("lookahead_words", "wordlm_cumsum", "multilevel_step"):
        monkeypatch.setattr(ops, name, getattr(ops_ref, name))
    sub, wrd = _dicts(gm)
    open_vocab, pen, weight = gm[variant + ".params"]
    m = MultiLevelLanguageModel(_lstm_lm(gm, "wsd.", "wlm_cfg", wrd), _lstm_lm(gm, "ssd.", "slm_cfg", sub), subwordlm_weight=float(weight),
                                oov_penalty=float(pen), open_vocab=bool(open_vocab))
    prev, orders, ref = (gm[variant + k] for k in (".prev_tokens", ".new_orders", ".out"))
    S, N = prev.shape
    tokens = torch.from_numpy(prev.T.copy()).to(torch.int32)
    state = m.init_incremental_state(None, N, 1)
    for t in range(S):
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32)
        out, is_logits = m.decode_step(t, tokens, state, order)
        assert not is_logits
        assert np.abs(out[:, : len(sub)].numpy() - ref[t]).max() < 2e-4, (variant, t)


def test_multilevel_closed_vocabulary_statements_agree(gm):
    """No reference output exists for open_vocab=False (upstream TypeError); the two oracle statements (pointer tree /
    CSR tree) of the evidently intended behaviour must at least agree with each other: dead hypotheses get logzero rows."""
    import torch

    from espresso_b200.tools.tensorized_prefix_tree import TensorizedPrefixTree
    from oracle import lookahead as OL
    from oracle import ops_ref

    sub, wrd = _dicts(gm)
    tree = TensorizedPrefixTree.build(wrd, sub)
    tr = {k: torch.from_numpy(getattr(tree, k)) for k in ("child_off", "child_tok", "child_node", "node_word")}
    root = OL.build_tree([wrd[i] for i in range(len(wrd))], {wrd.pad(), wrd.eos(), wrd.unk()}, sub.index, sub.unk())
    prev, orders, wl, sl = (gm["open" + k] for k in (".prev_tokens", ".new_orders", ".word_logprobs", ".sub_logprobs"))
    S, N = prev.shape
    Vs, Vw = len(sub), len(wrd)
    state = OL.MultiLevelState(root, N)
    nodes, nodes_tmp, words = torch.ones(N, dtype=torch.int32), torch.zeros(N, dtype=torch.int32), torch.zeros(N, dtype=torch.int32)
    wlp, out, cum = torch.zeros(N, Vw), torch.zeros(N, 24), torch.zeros(N)
    dead = 0
    for t in range(S):
        o = OL.multilevel_step(state, prev[t], wl[t], sl[t], t == 0, sub.space(), sub.eos(), wrd.unk(), wrd.eos(), 0.8, 1.0, False)
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32)
        if t > 0:
            ops_ref.lookahead_words(nodes, order, tr["node_word"], wrd.unk(), nodes_tmp, words)
        pt = torch.from_numpy(prev[t]).to(torch.int32)
        wlp_new, e = torch.zeros(N, Vw), torch.zeros(N)
        ops_ref.wordlm_cumsum(torch.from_numpy(wl[t]), Vw, pt, 1, sub.space(), t == 0, wlp, order, wlp_new, e, wrd.eos(), log_mode=True)
        wlp = wlp_new
        out_new, cum_new = torch.zeros(N, 24), torch.zeros(N)
        ops_ref.multilevel_step(pt, 1, t == 0, nodes if t == 0 else nodes_tmp, nodes, order, wlp, Vw, torch.from_numpy(sl[t]), False, 0.8,
                                out, cum, cum_new, tr, sub.space(), sub.eos(), wrd.unk(), wrd.eos(), 0.0, False, -10.0, out_new, Vs)
        out, cum = out_new, cum_new
        assert np.abs(out[:, :Vs].numpy() - o).max() < 1e-5, t
        dead += int((o[:, : sub.space()] == -10.0).all(-1).sum())
        if t < len(orders):
            state.reorder(orders[t])
    assert dead > 0


def test_wrap_language_models_follows_speech_recognize(gm):
    """espresso/speech_recognize.py:132-160: which fusion object a set of --lm-path models becomes."""
    from espresso_b200.models import MultiLevelLanguageModel, TensorizedLookaheadLanguageModel
    from espresso_b200.models.external_language_model import wrap_language_models

    sub, wrd = _dicts(gm)
    wlm, slm = _lstm_lm(gm, "wsd.", "wlm_cfg", wrd), _lstm_lm(gm, "ssd.", "slm_cfg", sub)
    wlm.is_wordlm = True
    assert wrap_language_models([slm], sub) is slm
    assert isinstance(wrap_language_models([wlm], sub), TensorizedLookaheadLanguageModel)
    ml = wrap_language_models([slm, wlm], sub, subwordlm_weight=0.5, oov_penalty=0.1)
    assert isinstance(ml, MultiLevelLanguageModel) and ml.subwordlm is slm and abs(ml.log_oov_penalty - np.log(0.1)) < 1e-12
    for bad in ([wlm, slm], [slm, slm], []):
        with pytest.raises(ValueError):
            wrap_language_models(bad, sub)
