"""GPU: beam-search kernels vs the oracle per-op reference, the reference's known-answer tests through the real
CUDA path, and bit-exact token agreement of the full generator with the oracle restatement of the reference
algorithm on randomised models (fp32 log-prob tables)."""
import sys
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_beam_merge_and_topk_vs_reference_ops():
    from espresso_b200 import ops
    from oracle import ops_ref as O

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for V, beam, bsz in ((50, 3, 2), (5004, 5, 4), (6, 2, 2)):
        N = bsz * beam
        ld = (V + 7) // 8 * 8
        x = (torch.randn(N, ld) * 3).bfloat16()
        lm = (torch.randn(N, ld) * 2).bfloat16()
        x[0, 5 % V] = float("nan")
        prev = torch.randn(N)
        for kw in (dict(), dict(force_eos=True), dict(eos_factor=1.5), dict(ban_eos=True, unk_penalty=0.7), dict(temperature=0.7)):
            out, outr = torch.empty(N, V, device=dev), torch.empty(N, V)
            ops.beam_merge(x.to(dev), V, True, out, prev_scores=prev.to(dev), lm=lm.to(dev), lm_weight=0.47, **kw)
            O.beam_merge(x, V, True, outr, prev_scores=prev, lm=lm, lm_weight=0.47, **kw)
            a, b = out.cpu(), outr
            assert torch.equal(torch.isinf(a), torch.isinf(b)), kw
            fin = ~torch.isinf(b)
            assert (a[fin] - b[fin]).abs().max().item() < 2e-4, kw
        # fp32 log-prob input, no LM
        lp = torch.log_softmax(torch.randn(N, V), -1)
        out, outr = torch.empty(N, V, device=dev), torch.empty(N, V)
        ops.beam_merge(lp.to(dev), V, False, out, prev_scores=None)
        O.beam_merge(lp, V, False, outr, prev_scores=None)
        fin = ~torch.isinf(outr)
        assert torch.equal(out.cpu()[fin], outr[fin])
        # top-k with ties: exact (score desc, index asc) order
        cand = torch.randn(bsz, beam * V)
        cand[:, ::7] = cand[:, 3:4]  # plenty of exact ties
        cand[0, :10] = float("-inf")
        K = min(2 * beam, beam * V - 1)
        s, t, b = ops.beam_topk(cand.to(dev), bsz, beam * V, beam * V, K, V)
        sr, tr, br = O.beam_topk(cand, bsz, beam * V, beam * V, K, V)
        assert torch.equal(s.cpu(), sr) and torch.equal(t.cpu(), tr) and torch.equal(b.cpu(), br)
        s, t, b = ops.beam_topk(cand.to(dev), bsz, beam * V, V, min(2 * beam, V - 1), V)  # step-0 form
        sr, tr, br = O.beam_topk(cand, bsz, beam * V, V, min(2 * beam, V - 1), V)
        assert torch.equal(s.cpu(), sr) and torch.equal(t.cpu(), tr) and torch.equal(b.cpu(), br)
    src = torch.randn(37, 24).bfloat16()
    idx = torch.randint(0, 37, (50,), dtype=torch.int32)
    assert torch.equal(ops.gather_rows(src.to(dev), idx.to(dev)).cpu(), src[idx.long()])


def test_reference_known_answers_on_gpu():
    from test_beam_search import BEAM_PROBS, CASES, EOS, V, W1, W2, _check, _Dict
    from espresso_b200.sequence_generator import SequenceGenerator, TableDecoderModel

    dev = torch.device("cuda:0")
    m = TableDecoderModel(BEAM_PROBS, V, EOS)
    sample = {"net_input": {"src_tokens": torch.tensor([[W1, W2, EOS], [W1, W2, EOS]], device=dev),
                            "src_lengths": torch.tensor([2, 2], device=dev)}}
    for name, (kw, chk, exp) in CASES.items():
        hyp = SequenceGenerator([m], _Dict(), beam_size=2, **kw).generate([m], sample)
        flat = [hyp[0][0], hyp[0][1], hyp[1][0], hyp[1][1]]
        for h, (t, p) in zip(flat, exp):
            _check(h, t, p, **chk)


@pytest.mark.parametrize("seed,beam,bsz,Vn,eos_factor", [(1, 2, 3, 9, None), (2, 5, 4, 17, 1.5), (5, 5, 6, 64, 1.5), (6, 8, 2, 33, None)])
def test_generator_tokens_bit_exact_vs_oracle(seed, beam, bsz, Vn, eos_factor):
    from test_beam_search import EOS, PAD, UNK, _Dict, _RandomModel
    from espresso_b200.sequence_generator import SequenceGenerator
    from oracle import beam as OB

    dev = torch.device("cuda:0")

    class D(_Dict):
        def __len__(self):
            return Vn

    class GpuModel(_RandomModel):
        def decode_step(self, step, tokens, state, new_order):
            return self.lprobs(step, tokens.cpu()).to(dev), False

    m = GpuModel(Vn, seed)
    kw = dict(beam_size=beam, max_len_a=0.0, max_len_b=14, min_len=2, unk_penalty=0.3, eos_factor=eos_factor)
    sample = {"net_input": {"src_tokens": torch.zeros(bsz, 7, dtype=torch.long, device=dev), "src_lengths": torch.full((bsz,), 7, device=dev)}}
    got = SequenceGenerator([m], D(), **kw).generate([m], sample)
    ref = OB.generate(lambda step, tokens, ro: m.lprobs(step, tokens), bsz, 7, Vn, PAD, UNK, EOS, model_max_len=m.max_pos, **kw)
    for hs, rs in zip(got, ref):
        assert len(hs) == len(rs)
        for h, r in zip(hs, rs):
            assert h["tokens"].tolist() == r["tokens"].tolist()      # bit-exact token indices
            assert abs(float(h["score"]) - float(r["score"])) < 1e-5  # scores within 1e-5
