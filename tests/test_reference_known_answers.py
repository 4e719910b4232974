"""The known-answer vectors and properties of the reference's OWN unit tests for components on the path, replayed against
this repository's implementations:
  tests/espresso/test_speech_utils.py:127-163   collate_frames
  tests/espresso/test_speech_utils.py:221-263   edit_distance counters
  tests/test_label_smoothing.py:62-127          label-smoothed CE: nll == plain CE, padding ignored, per-row reduction,
                                                zero smoothing == CE
(the beam-search tables of tests/test_sequence_generator.py are in tests/test_beam_search.py)."""
import os

import pytest
import torch

from oracle import ops_ref as O


def test_collate_frames_reference_vectors():
    from espresso_b200.data.collate import collate_frames

    vals = [torch.tensor([4.5, 2.3, 1.2]).unsqueeze(-1).expand(-1, 10), torch.tensor([6.7, 9.8]).unsqueeze(-1).expand(-1, 10),
            torch.tensor([7.7, 5.4, 6.2, 8.0]).unsqueeze(-1).expand(-1, 10), torch.tensor([1.5]).unsqueeze(-1).expand(-1, 10)]
    expected = torch.tensor([[4.5, 2.3, 1.2, 0.0], [6.7, 9.8, 0.0, 0.0], [7.7, 5.4, 6.2, 8.0], [1.5, 0.0, 0.0, 0.0]]).unsqueeze(-1).expand(-1, -1, 10)
    assert torch.equal(collate_frames(vals, pad_value=0.0), expected)


def test_edit_distance_reference_counters():
    from espresso_b200.tasks.speech_recognition import edit_counts

    # (ref, hyp) -> sub + ins + del, words   from the reference's expected Counters
    cases = [([], [], 0, 0), (["a", "b", "c"], [], 3, 3), (["a", "b", "c"], ["a", "b", "c"], 0, 3),
             (["a", "b", "c"], ["d", "b", "c", "e", "f"], 3, 3), (["b", "c", "d", "e", "f", "h"], ["d", "b", "c", "e", "f", "g"], 3, 6)]
    for ref, hyp, err, words in cases:
        assert edit_counts(ref, hyp) == (err, words)


def _toy():
    torch.manual_seed(0)
    V, pad = 11, 1
    logits = (torch.randn(2, 4, 16) * 1.5).bfloat16()           # 16-wide rows, 11 valid classes
    target = torch.tensor([[4, 5, 6, 2], [7, 2, pad, pad]])
    return logits, target, V, pad


def test_label_smoothing_properties_of_the_reference_tests():
    logits, target, V, pad = _toy()
    x, t = logits.view(-1, 16), target.view(-1).int()
    plain = torch.nn.functional.cross_entropy(x.float()[:, :V], target.view(-1), ignore_index=pad, reduction="none")
    # test_nll_loss: the nll part of the smoothed criterion is plain cross-entropy
    loss, nll, _ = O.lsce_loss(x, V, t, pad, 0.1)
    assert torch.allclose(nll, plain, atol=1e-5)
    # test_zero_eps: no smoothing -> the loss itself is cross-entropy
    loss0, nll0, _ = O.lsce_loss(x, V, t, pad, 0.0)
    assert torch.allclose(loss0, plain, atol=1e-5)
    # test_padding: a padded batch scores like its sentences scored alone
    l_a, _, _ = O.lsce_loss(logits[0], V, target[0].int(), pad, 0.1)
    l_b, _, _ = O.lsce_loss(logits[1, :2], V, target[1, :2].int(), pad, 0.1)
    assert abs(float(loss.sum()) - float(l_a.sum() + l_b.sum())) < 1e-5
    # test_reduction: the reduced loss is the sum of the per-token losses, padding contributes zero
    assert float(loss.view(2, 4)[1, 2:].abs().sum()) == 0.0


def test_wer_scorer_replays_the_reference(golden_dir, tmp_path):
    """espresso_b200.tools.wer.Scorer on 60 recorded utterances (non-linguistic symbols, sed-style word filters, reversed print
    order): edit counters, CER / WER with their shares and the aligned REF / HYP / STP blocks equal what the REAL reference
    Scorer produced (tests/golden/wer_scorer.json, made by oracle/pin_against_reference.py wer_scorer)."""
    import json

    from espresso_b200.tools.wer import Scorer, align

    g = json.load(open(os.path.join(golden_dir, "wer_scorer.json"), encoding="utf-8"))

    class _D:
        non_lang_syms = ["<noise>", "<laugh>"]

        @staticmethod
        def wordpiece_decode(x):
            return x.replace(" ", "").replace("▁", " ").strip()

    filt = tmp_path / "words.filt"
    filt.write_text(g["filter"], encoding="utf-8")
    sc = Scorer(_D(), wer_output_filter=str(filt))
    for uid, ref, hyp in g["utts"]:
        sc.add_prediction(uid, hyp)
        sc.add_evaluation(uid, ref, hyp)
    sc.add_ordered_utt_list(g["order"])
    assert dict(sc.char_counter) == g["char_counter"] and dict(sc.word_counter) == g["word_counter"]
    assert list(sc.cer()) == g["cer"] and list(sc.wer()) == g["wer"]
    assert sc.print_char_results() == g["print_char_results"]
    assert sc.print_results() == g["print_results"]
    assert sc.print_aligned_results() == g["print_aligned_results"]
    assert sc.tot_word_error() == sum(g["word_counter"][k] for k in ("sub", "ins", "del")) and sc.tot_word_count() == g["word_counter"]["words"]
    with pytest.raises(AssertionError):
        sc.add_prediction(g["utts"][0][0], "x")  # duplicated utterance id
    with pytest.raises(TypeError):
        sc.add_evaluation(3, "a", "b")
    # edge cases of the alignment itself
    assert align([], []) == [] and align(["a"], []) == ["del"] and align([], ["a", "b"]) == ["ins", "ins"]
    assert align(["a", "b", "c"], ["a", "x", "c"]) == ["corr", "sub", "corr"]
