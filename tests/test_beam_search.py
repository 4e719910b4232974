"""CPU: the reference's own known-answer beam-search tests (tests/test_sequence_generator.py:202-283, probability
tables from tests/utils.py:67-166) replayed against (1) the oracle restatement of the reference algorithm and
(2) the product SequenceGenerator with the oracle's per-op kernels monkeypatched in; plus randomised equivalence
of the two (tokens bit-exact).  The CUDA kernels are checked against the same per-op reference in test_gpu_beam.py."""
import os
import math

import numpy as np
import pytest
import torch

from oracle import beam as OB
from oracle import ops_ref

PAD, EOS, UNK, W1, W2, V = 1, 2, 3, 4, 5, 6  # fairseq Dictionary: <s> <pad> </s> <unk> token_0 token_1
U = 0.0
BEAM_PROBS = [
    torch.tensor([[0.0, U, 0.9, 0.1], [0.0, U, 0.9, 0.1], [0.0, U, 0.7, 0.3], [0.0, U, 0.7, 0.3]]),
    torch.tensor([[1.0, U, 0.0, 0.0], [0.0, U, 0.9, 0.1], [0.25, U, 0.35, 0.4], [0.00, U, 0.10, 0.9]]),
    torch.tensor([[0.0, U, 0.1, 0.9], [0.6, U, 0.2, 0.2], [0.60, U, 0.4, 0.00], [0.01, U, 0.0, 0.99]]),
    torch.tensor([[1.0, U, 0.0, 0.0], [1.0, U, 0.0, 0.0], [0.1, U, 0.5, 0.4], [1.0, U, 0.0, 0.0]]),
]


class _Dict:
    def __len__(self):
        return V

    def pad(self):
        return PAD

    def eos(self):
        return EOS

    def unk(self):
        return UNK


@pytest.fixture()
def cpu_ops(monkeypatch):
    from espresso_b200 import ops

    for name in dir(ops_ref):
        if name.startswith("_") or not callable(getattr(ops_ref, name)) or not hasattr(ops, name):
            continue
        monkeypatch.setattr(ops, name, getattr(ops_ref, name))
    return ops


def _table_fn(step, tokens, reorder):
    rows = tokens.shape[0]
    probs = torch.zeros(rows, V)
    if step < len(BEAM_PROBS):
        probs[:, EOS:] = BEAM_PROBS[step][:rows] if BEAM_PROBS[step].shape[0] >= rows else BEAM_PROBS[step]
    else:
        probs[:, EOS] = 1.0
    return probs.log()


def _oracle(**kw):
    return OB.generate(_table_fn, 2, 3, V, PAD, UNK, EOS, beam_size=2, model_max_len=100, **kw)


def _product(cpu_ops, **kw):
    from espresso_b200.sequence_generator import SequenceGenerator, TableDecoderModel

    m = TableDecoderModel(BEAM_PROBS, V, EOS)
    g = SequenceGenerator([m], _Dict(), beam_size=2, **kw)
    sample = {"net_input": {"src_tokens": torch.tensor([[W1, W2, EOS], [W1, W2, EOS]]), "src_lengths": torch.tensor([2, 2])}}
    return g.generate([m], sample)


def _check(h, toks, probs, normalized=True, lenpen=1.0):
    assert h["tokens"].tolist() == toks
    ps = torch.tensor(probs).log()
    assert (h["positional_scores"] - ps).abs().max() < 1e-4
    sc = ps.sum() / (len(probs) ** lenpen if normalized else 1.0)
    assert abs(float(sc) - float(h["score"])) < 1e-5


CASES = {
    "with_normalization": (dict(), dict(), [
        ([W1, EOS], [0.9, 1.0]), ([W2, W1, W2, EOS], [0.1, 0.9, 0.9, 1.0]),
        ([W1, W2, W1, EOS], [0.7, 0.4, 0.4, 1.0]), ([W1, W2, EOS], [0.7, 0.4, 0.6])]),
    "without_normalization": (dict(normalize_scores=False), dict(normalized=False), [
        ([W1, EOS], [0.9, 1.0]), ([W2, W1, W2, EOS], [0.1, 0.9, 0.9, 1.0]),
        ([W1, W2, EOS], [0.7, 0.4, 0.6]), ([W1, W2, W1, EOS], [0.7, 0.4, 0.4, 1.0])]),
    "lenpen_short": (dict(len_penalty=0.6), dict(lenpen=0.6), [
        ([W1, EOS], [0.9, 1.0]), ([W2, W1, W2, EOS], [0.1, 0.9, 0.9, 1.0]),
        ([W1, W2, EOS], [0.7, 0.4, 0.6]), ([W1, W2, W1, EOS], [0.7, 0.4, 0.4, 1.0])]),
    "lenpen_long": (dict(len_penalty=5.0), dict(lenpen=5.0), [
        ([W2, W1, W2, EOS], [0.1, 0.9, 0.9, 1.0]), ([W1, EOS], [0.9, 1.0]),
        ([W1, W2, W1, EOS], [0.7, 0.4, 0.4, 1.0]), ([W1, W2, EOS], [0.7, 0.4, 0.6])]),
    "maxlen": (dict(max_len_b=2), dict(), [
        ([W1, EOS], [0.9, 1.0]), ([W2, W2, EOS], [0.1, 0.1, 0.6]),
        ([W1, W2, EOS], [0.7, 0.4, 0.6]), ([W2, W2, EOS], [0.3, 0.9, 0.01])]),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_known_answers_oracle(case):
    kw, chk, exp = CASES[case]
    hyp = _oracle(**kw)
    flat = [hyp[0][0], hyp[0][1], hyp[1][0], hyp[1][1]]
    for h, (t, p) in zip(flat, exp):
        _check(h, t, p, **chk)


@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_known_answers_product(case, cpu_ops):
    kw, chk, exp = CASES[case]
    hyp = _product(cpu_ops, **kw)
    flat = [hyp[0][0], hyp[0][1], hyp[1][0], hyp[1][1]]
    for h, (t, p) in zip(flat, exp):
        _check(h, t, p, **chk)


class _RandomModel:
    """Deterministic pseudo-random log-prob tables that depend on the hypothesis prefix (so reordering matters)."""

    def __init__(self, Vn, seed, max_pos=40):
        self.V, self.seed, self.max_pos = Vn, seed, max_pos

    def max_decoder_positions(self):
        return self.max_pos

    def forward_encoder(self, net_input):
        return None

    def init_incremental_state(self, enc, bsz, beam):
        return {}

    def lprobs(self, step, tokens):
        rows = tokens.shape[0]
        out = torch.empty(rows, self.V)
        for r in range(rows):
            h = hash((self.seed, step) + tuple(int(t) for t in tokens[r, : step + 1].tolist())) % (2 ** 31)
            g = torch.Generator().manual_seed(h)
            out[r] = torch.log_softmax(torch.randn(self.V, generator=g) * 2.0, dim=-1)
        return out

    def decode_step(self, step, tokens, state, new_order):
        return self.lprobs(step, tokens), False


@pytest.mark.parametrize("seed,beam,bsz,Vn,eos_factor,lenpen", [(1, 2, 3, 9, None, 1.0), (2, 5, 4, 17, 1.5, 1.0), (3, 3, 2, 8, None, 0.5),
                                                              (4, 4, 3, 6, 2.0, 1.0)])
def test_product_matches_oracle_on_random_models(seed, beam, bsz, Vn, eos_factor, lenpen, cpu_ops):
    from espresso_b200.sequence_generator import SequenceGenerator

    class D(_Dict):
        def __len__(self):
            return Vn

    m = _RandomModel(Vn, seed)
    kw = dict(beam_size=beam, max_len_a=0.0, max_len_b=12, min_len=2, len_penalty=lenpen, unk_penalty=0.3, eos_factor=eos_factor)
    g = SequenceGenerator([m], D(), **kw)
    sample = {"net_input": {"src_tokens": torch.zeros(bsz, 7, dtype=torch.long), "src_lengths": torch.full((bsz,), 7)}}
    got = g.generate([m], sample)
    ref = OB.generate(lambda step, tokens, ro: m.lprobs(step, tokens), bsz, 7, Vn, PAD, UNK, EOS, model_max_len=m.max_pos, **kw)
    assert len(got) == len(ref) == bsz
    for hs, rs in zip(got, ref):
        assert len(hs) == len(rs)
        for h, r in zip(hs, rs):
            assert h["tokens"].tolist() == r["tokens"].tolist()
            assert abs(float(h["score"]) - float(r["score"])) < 1e-5
            assert (h["positional_scores"] - r["positional_scores"]).abs().max() < 1e-5


# ---------------------------------------------------------------------------------------------------
# hypotheses of the REAL reference SequenceGenerator (LM shallow fusion, eos_factor, unk penalty, min_len, length
# penalty) recorded by oracle/pin_against_reference.py::pin_beam -> tests/golden/beam_reference.npz
# ---------------------------------------------------------------------------------------------------
def _reference_cases(golden_dir):
    import numpy as np

    g = np.load(os.path.join(golden_dir, "beam_reference.npz"))
    for ci, (seed, beam, bsz, Vn, eos_factor, lenpen, lm) in enumerate(g["cases"].tolist()):
        ref = [[(g["c%d_b%d_k%d_tokens" % (ci, b, k)].tolist(), float(g["c%d_b%d_k%d_score" % (ci, b, k)]))
                for k in range(int(g["c%d_b%d_n" % (ci, b)]))] for b in range(int(bsz))]
        yield dict(seed=int(seed), beam=int(beam), bsz=int(bsz), V=int(Vn), eos_factor=None if eos_factor < 0 else eos_factor,
                   lenpen=lenpen, lm=None if lm < 0 else lm), ref


def test_oracle_beam_search_matches_reference_generator(golden_dir):
    n = 0
    for c, ref in _reference_cases(golden_dir):
        m, lm = _RandomModel(c["V"], c["seed"]), _RandomModel(c["V"], c["seed"] + 1000)

        def fn(step, tokens, ro):
            lp = m.lprobs(step, tokens)
            return lp if c["lm"] is None else lp + c["lm"] * lm.lprobs(step, tokens)

        got = OB.generate(fn, c["bsz"], 7, c["V"], PAD, UNK, EOS, model_max_len=40, beam_size=c["beam"], max_len_a=0.0,
                          max_len_b=12, min_len=2, len_penalty=c["lenpen"], unk_penalty=0.3, eos_factor=c["eos_factor"])
        for hs, rs in zip(got, ref):
            assert len(hs) == len(rs)
            for h, (toks, score) in zip(hs, rs):
                assert h["tokens"].tolist() == toks and abs(float(h["score"]) - score) < 1e-5
                n += 1
    assert n > 50


def test_product_generator_matches_reference_generator(golden_dir, cpu_ops):
    """espresso_b200.SequenceGenerator (host orchestration over the reference ops) reproduces the real reference
    generator's hypotheses, token for token, incl. LM shallow fusion and eos_factor."""
    from espresso_b200.sequence_generator import SequenceGenerator

    for c, ref in _reference_cases(golden_dir):
        Vn = c["V"]

        class D(_Dict):
            def __len__(self):
                return Vn

        m = _RandomModel(Vn, c["seed"])
        lm = _RandomModel(Vn, c["seed"] + 1000) if c["lm"] is not None else None
        gen = SequenceGenerator([m], D(), beam_size=c["beam"], max_len_a=0.0, max_len_b=12, min_len=2, len_penalty=c["lenpen"],
                                unk_penalty=0.3, eos_factor=c["eos_factor"], lm_model=lm, lm_weight=c["lm"] or 1.0)
        sample = {"net_input": {"src_tokens": torch.zeros(c["bsz"], 7, dtype=torch.long), "src_lengths": torch.full((c["bsz"],), 7)}}
        got = gen.generate([m], sample)
        for hs, rs in zip(got, ref):
            assert len(hs) == len(rs)
            for h, (toks, score) in zip(hs, rs):
                assert h["tokens"].tolist() == toks, (c, toks, h["tokens"].tolist())
                assert abs(float(h["score"]) - score) < 1e-4


def test_ensemble_averages_probabilities_like_the_reference(cpu_ops):
    """Two models: per-step log-probs are averaged in the probability domain (logsumexp - log N,
    fairseq/sequence_generator.py:895-897) before LM fusion and the search; checked against the oracle driven with the
    averaged tables."""
    from espresso_b200.sequence_generator import SequenceGenerator

    Vn, bsz, beam = 17, 3, 4

    class D(_Dict):
        def __len__(self):
            return Vn

    m1, m2, lm = _RandomModel(Vn, 21), _RandomModel(Vn, 22), _RandomModel(Vn, 23)
    kw = dict(beam_size=beam, max_len_a=0.0, max_len_b=10, min_len=2, unk_penalty=0.2, eos_factor=1.5)
    gen = SequenceGenerator([m1, m2], D(), lm_model=lm, lm_weight=0.4, **kw)
    sample = {"net_input": {"src_tokens": torch.zeros(bsz, 7, dtype=torch.long), "src_lengths": torch.full((bsz,), 7)}}
    got = gen.generate([m1, m2], sample)

    def fn(step, tokens, ro):
        avg = torch.logsumexp(torch.stack([m1.lprobs(step, tokens), m2.lprobs(step, tokens)]), dim=0) - math.log(2)
        return avg + 0.4 * lm.lprobs(step, tokens)

    ref = OB.generate(fn, bsz, 7, Vn, PAD, UNK, EOS, model_max_len=m1.max_pos, **kw)
    for hs, rs in zip(got, ref):
        assert len(hs) == len(rs)
        for h, r in zip(hs, rs):
            assert h["tokens"].tolist() == r["tokens"].tolist() and abs(float(h["score"]) - float(r["score"])) < 1e-4
