"""GPU: look-ahead word-LM fusion kernels (csrc/lookahead.cu) and the TensorizedLookaheadLanguageModel wrapper against
outputs recorded from the REAL reference class (tests/golden/lookahead_lm.npz) and, inside the beam search, against the
oracle restatement.  Tolerance: 2e-4 on log-probabilities above -15 (probabilities are ratios of differences of fp32
cumulative sums; below 3e-7 both sides are cancellation noise, bounded by 1.0 in the log domain)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [("open", True, 1e-4), ("closed", False, 1e-4), ("open_pen", True, 0.3)]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "lookahead_lm.npz"))


def _setup(g):
    from test_lookahead_lm import _dicts

    return _dicts(g)


def _close(got, ref, what):
    big = ref > -15
    d = np.abs(got - ref)
    assert d.max() < 1.0, (what, d.max())
    if big.any():
        assert d[big].max() < 2e-4, (what, d[big].max())


def _word_lm(g, wrd, dev, dtype=torch.float32):
    from espresso_b200.models import LSTMLanguageModelEspresso, LSTMLanguageModelEspressoConfig

    class _Task:
        target_dictionary = wrd

    e, h, o, nl = (int(v) for v in g["lm_cfg"])
    lm = LSTMLanguageModelEspresso.build_model(LSTMLanguageModelEspressoConfig(
        dropout=0.0, decoder_embed_dim=e, decoder_hidden_size=h, decoder_layers=nl, decoder_out_embed_dim=o), _Task())
    lm.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    return lm.finalize_(dev, dtype) if dev is not None else lm.eval()


@pytest.mark.parametrize("variant,open_vocab,pen", VARIANTS)
def test_kernels_vs_reference_fixture(g, variant, open_vocab, pen):
    """The three launches of one search step, fed with the reference LM's recorded word distributions."""
    from espresso_b200 import ops
    from espresso_b200.tools.tensorized_prefix_tree import TensorizedPrefixTree

    dev = torch.device("cuda:0")
    sub, wrd = _setup(g)
    tree = TensorizedPrefixTree.build(wrd, sub)
    tr = tree.to(dev)
    prev, orders, ref, lm = (g[variant + k] for k in (".prev_tokens", ".new_orders", ".out", ".lm_probs"))
    S, N = prev.shape
    Vs, Vw = len(sub), len(wrd)
    tokens = torch.from_numpy(prev.T.copy()).to(torch.int32).to(dev)          # [N, S]: column t = previous subword at step t
    nodes = torch.full((N,), 1, dtype=torch.int32, device=dev)
    nodes_tmp, words = torch.zeros_like(nodes), torch.zeros_like(nodes)
    cum, cum_alt = torch.zeros(N, Vw, device=dev), torch.zeros(N, Vw, device=dev)
    eos_lp = torch.zeros(N, device=dev)
    out = torch.zeros(N, (Vs + 7) // 8 * 8, device=dev)
    for t in range(S):
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32).to(dev)
        if t > 0:
            ops.lookahead_words(nodes, order, tr["node_word"], wrd.unk(), nodes_tmp, words)
        logits = torch.from_numpy(np.log(np.maximum(lm[t], 1e-30))).to(dev)   # softmax(log p) = p
        ops.wordlm_cumsum(logits, Vw, tokens[:, t], tokens.stride(0), sub.space(), t == 0, cum, order, cum_alt, eos_lp, wrd.eos())
        cum, cum_alt = cum_alt, cum
        ops.lookahead_step(tokens[:, t], tokens.stride(0), t == 0, nodes if t == 0 else nodes_tmp, nodes, cum, Vw, eos_lp, tr,
                           sub.space(), sub.eos(), sub.pad(), wrd.unk(), pen, open_vocab, 1e-10, out, Vs)
        torch.cuda.synchronize()
        _close(out[:, :Vs].cpu().numpy(), ref[t], (variant, t))
        assert torch.isinf(out[:, Vs:]).all()
        # cumulative sums are monotone and end at 1
        c = cum.cpu().numpy()
        assert (np.diff(c, axis=1) >= -1e-7).all() and np.abs(c[:, -1] - 1).max() < 1e-5
    # bf16 logits take the same path (the word LM runs in bf16 in production): a looser, distribution-level check
    lg = torch.randn(N, Vw, device=dev) * 3
    c32, c16, e = torch.empty(N, Vw, device=dev), torch.empty(N, Vw, device=dev), torch.empty(N, device=dev)
    ops.wordlm_cumsum(lg, Vw, tokens[:, 0], tokens.stride(0), sub.space(), True, cum, None, c32, e, wrd.eos())
    ops.wordlm_cumsum(lg.bfloat16(), Vw, tokens[:, 0], tokens.stride(0), sub.space(), True, cum, None, c16, e, wrd.eos())
    want = torch.softmax(lg.double(), -1).cumsum(-1)
    assert (c32.double() - want).abs().max().item() < 2e-6 and (c16.double() - want).abs().max().item() < 3e-2
    assert (e - torch.log_softmax(lg.bfloat16().float(), -1)[:, wrd.eos()]).abs().max().item() < 1e-4


def test_cumsum_large_vocabulary():
    """|words| = 65 003 (a LibriSpeech-size word LM), 250 hypotheses: fp32 output within 2e-6 of a float64 scan."""
    from espresso_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    N, Vw = 250, 65003
    lg = torch.randn(N, Vw + 5, device=dev)[:, :Vw] * 4          # row stride != Vw
    prev = torch.zeros(N, 1, dtype=torch.int32, device=dev)
    prev[::3] = 7                                                # "space": these rows are rescanned, the others inherit
    old = torch.rand(N, Vw, device=dev)
    order = torch.randint(0, N, (N,), dtype=torch.int32, device=dev)
    out, e = torch.empty(N, Vw, device=dev), torch.empty(N, device=dev)
    ops.wordlm_cumsum(lg, Vw, prev[:, 0], 1, 7, False, old, order, out, e, 2)
    want = torch.softmax(lg.double(), -1).cumsum(-1)
    fresh = (prev[:, 0] == 7)
    assert (out[fresh].double() - want[fresh]).abs().max().item() < 2e-6
    assert torch.equal(out[~fresh], old[order.long()][~fresh])
    assert (e[fresh].double() - torch.log_softmax(lg.double(), -1)[fresh, 2]).abs().max().item() < 1e-4


@pytest.mark.parametrize("variant,open_vocab,pen", VARIANTS)
def test_model_vs_reference_fixture(g, variant, open_vocab, pen):
    """LSTM word LM (fixture weights, fp32) + wrapper, driven like the generator drives it."""
    from espresso_b200.models import TensorizedLookaheadLanguageModel

    dev = torch.device("cuda:0")
    sub, wrd = _setup(g)
    m = TensorizedLookaheadLanguageModel(_word_lm(g, wrd, None), sub, oov_penalty=pen, open_vocab=open_vocab).finalize_(dev, torch.float32)
    prev, orders, ref = (g[variant + k] for k in (".prev_tokens", ".new_orders", ".out"))
    S, N = prev.shape
    tokens = torch.from_numpy(prev.T.copy()).to(torch.int32).to(dev)
    state = m.init_incremental_state(None, N, 1)
    for t in range(S):
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32).to(dev)
        out, is_logits = m.decode_step(t, tokens, state, order)
        assert not is_logits
        _close(out[:, : len(sub)].cpu().numpy(), ref[t], (variant, t))


def test_beam_search_with_lookahead_fusion_matches_oracle(g):
    """SequenceGenerator(lm_model=look-ahead LM) on the GPU vs the oracle beam search whose step log-probs are
    acoustic table + lm_weight * oracle look-ahead (word distributions from the same LSTM LM on the CPU)."""
    from test_beam_search import _RandomModel
    from espresso_b200.models import TensorizedLookaheadLanguageModel
    from espresso_b200.sequence_generator import SequenceGenerator
    from oracle import beam as OB
    from oracle import lookahead as OL

    dev = torch.device("cuda:0")
    sub, wrd = _setup(g)
    Vs, beam, bsz, lmw = len(sub), 4, 3, 0.6

    class GpuModel(_RandomModel):
        def decode_step(self, step, tokens, state, new_order):
            return self.lprobs(step, tokens.cpu()).to(dev), False

    am = GpuModel(Vs, 11)
    lm_gpu = TensorizedLookaheadLanguageModel(_word_lm(g, wrd, None), sub).finalize_(dev, torch.float32)
    kw = dict(beam_size=beam, max_len_a=0.0, max_len_b=12, min_len=1, eos_factor=None)
    sample = {"net_input": {"src_tokens": torch.zeros(bsz, 7, dtype=torch.long, device=dev), "src_lengths": torch.full((bsz,), 7, device=dev)}}
    got = SequenceGenerator([am], sub, lm_model=lm_gpu, lm_weight=lmw, **kw).generate([am], sample)

    lm_cpu = _word_lm(g, wrd, None).float()
    dec = lm_cpu.decoder
    root = OL.build_tree([wrd[i] for i in range(len(wrd))], {wrd.pad(), wrd.eos(), wrd.unk()}, sub.index, sub.unk())
    st = {"la": OL.LookaheadState(root, bsz * beam), "h": None, "c": None}

    def lprobs_fn(step, tokens, reorder):
        N = tokens.shape[0]
        prev = tokens[:, step].numpy()
        with torch.no_grad():
            if step == 0:
                w = torch.full((N,), wrd.eos())
                h = [torch.zeros(N, dec.hidden_size) for _ in dec.layers]
                c = [torch.zeros(N, dec.hidden_size) for _ in dec.layers]
            else:
                if reorder is not None:
                    st["la"].reorder(reorder.tolist())
                    st["h"] = [x[reorder] for x in st["h"]]
                    st["c"] = [x[reorder] for x in st["c"]]
                w = torch.from_numpy(st["la"].lm_words(wrd.unk()))
                h, c = st["h"], st["c"]
            y, h2, c2, _ = dec.step(dec.embed_tokens(w), h, c, None)
            probs = torch.softmax(dec.output_layer(y).float(), -1).numpy()
            keep = torch.from_numpy(prev == sub.space())[:, None] if step > 0 else torch.ones(N, 1, dtype=torch.bool)
            st["h"] = [torch.where(keep, a, b) for a, b in zip(h2, h)]
            st["c"] = [torch.where(keep, a, b) for a, b in zip(c2, c)]
        lm_lp = OL.step(st["la"], prev, probs, step == 0, Vs, sub.space(), sub.eos(), sub.pad(), wrd.unk(), wrd.eos())
        return am.lprobs(step, tokens) + lmw * torch.from_numpy(lm_lp).float()

    ref = OB.generate(lprobs_fn, bsz, 7, Vs, sub.pad(), sub.unk(), sub.eos(), model_max_len=am.max_pos, **kw)
    n_tok = 0
    for hs, rs in zip(got, ref):
        assert len(hs) == len(rs)
        for h, r in zip(hs, rs):
            assert h["tokens"].tolist() == r["tokens"].tolist()
            assert abs(float(h["score"]) - float(r["score"])) < 1e-4
            n_tok += len(r["tokens"])
    assert n_tok > bsz * beam * 2


# ---- multi-level (subword + word) LM -----------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gm(golden_dir):
    return np.load(os.path.join(golden_dir, "multilevel_lm.npz"))


@pytest.mark.parametrize("variant", ["open", "open_pen", "open_w1"])
def test_multilevel_model_vs_reference_fixture(gm, variant):
    """Both LSTM LMs with the fixture's weights (fp32) + the wrapper, through esp_wordlm_cumsum(log mode) and
    esp_multilevel_step, against rows recorded from the real MultiLevelLanguageModel."""
    from test_lookahead_lm import _dicts, _lstm_lm
    from espresso_b200.models import MultiLevelLanguageModel

    dev = torch.device("cuda:0")
    sub, wrd = _dicts(gm)
    open_vocab, pen, weight = gm[variant + ".params"]
    m = MultiLevelLanguageModel(_lstm_lm(gm, "wsd.", "wlm_cfg", wrd), _lstm_lm(gm, "ssd.", "slm_cfg", sub), subwordlm_weight=float(weight),
                                oov_penalty=float(pen), open_vocab=bool(open_vocab)).finalize_(dev, torch.float32)
    prev, orders, ref = (gm[variant + k] for k in (".prev_tokens", ".new_orders", ".out"))
    S, N = prev.shape
    tokens = torch.from_numpy(prev.T.copy()).to(torch.int32).to(dev)
    state = m.init_incremental_state(None, N, 1)
    for t in range(S):
        order = None if t == 0 else torch.from_numpy(orders[t - 1]).to(torch.int32).to(dev)
        out, is_logits = m.decode_step(t, tokens, state, order)
        assert not is_logits
        assert np.abs(out[:, : len(sub)].cpu().numpy() - ref[t]).max() < 2e-4, (variant, t)
        assert torch.isinf(out[:, len(sub):]).all()


@pytest.mark.parametrize("open_vocab,is_logits,dtype", [(True, True, torch.bfloat16), (False, False, torch.float32), (False, True, torch.float32)])
def test_multilevel_kernel_vs_oracle_ops(gm, open_vocab, is_logits, dtype):
    """esp_multilevel_step vs oracle/ops_ref.multilevel_step on random rows, incl. the closed-vocabulary branch the
    reference cannot run, bf16 logits and a 5 004-unit subword vocabulary width."""
    from test_lookahead_lm import _dicts
    from espresso_b200 import ops
    from espresso_b200.tools.tensorized_prefix_tree import TensorizedPrefixTree
    from oracle import ops_ref

    dev = torch.device("cuda:0")
    sub, wrd = _dicts(gm)
    tree = TensorizedPrefixTree.build(wrd, sub)
    tr_cpu = {k: torch.from_numpy(getattr(tree, k)) for k in ("child_off", "child_tok", "child_node", "node_word")}
    tr = tree.to(dev)
    torch.manual_seed(1)
    N, Vw, Vs = 37, len(wrd), len(sub)
    ld = (Vs + 7) // 8 * 8
    for first in (True, False):
        prev = torch.randint(0, Vs, (N,), dtype=torch.int32)
        prev[::4] = sub.space()
        prev[1::9] = sub.eos()
        nodes_in = torch.randint(0, tree.num_nodes, (N,), dtype=torch.int32)
        order = torch.randint(0, N, (N,), dtype=torch.int32)
        wlp = torch.log_softmax(torch.randn(N, Vw), -1)
        x = torch.randn(N, ld) * 2
        if not is_logits:
            x[:, :Vs] = torch.log_softmax(x[:, :Vs], -1)
        x = x.to(dtype)
        out_prev, cum_in = torch.randn(N, ld), torch.randn(N)
        args = (sub.space(), sub.eos(), wrd.unk(), wrd.eos(), -1.7, open_vocab, -10.0)
        o_ref, c_ref, n_ref = torch.zeros(N, ld), torch.zeros(N), torch.zeros(N, dtype=torch.int32)
        ops_ref.multilevel_step(prev, 1, first, nodes_in, n_ref, order, wlp, Vw, x.float(), is_logits, 0.7, out_prev, cum_in, c_ref, tr_cpu,
                                *args, o_ref, Vs)
        o, c, n = torch.zeros(N, ld, device=dev), torch.zeros(N, device=dev), torch.zeros(N, dtype=torch.int32, device=dev)
        ops.multilevel_step(prev.to(dev), 1, first, nodes_in.to(dev), n, order.to(dev), wlp.to(dev), Vw, x.to(dev), is_logits, 0.7,
                            out_prev.to(dev), cum_in.to(dev), c, tr, *args, o, Vs)
        assert torch.equal(n.cpu(), n_ref)
        assert (c.cpu() - c_ref).abs().max().item() < 1e-5
        assert (o.cpu()[:, :Vs] - o_ref[:, :Vs]).abs().max().item() < (2e-2 if dtype == torch.bfloat16 else 1e-4)
