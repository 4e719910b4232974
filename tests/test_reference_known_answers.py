"""The known-answer vectors and properties of the reference's OWN unit tests for components on the path, replayed against
this repository's implementations:
  tests/espresso/test_speech_utils.py:127-163   collate_frames
  tests/espresso/test_speech_utils.py:221-263   edit_distance counters
  tests/test_label_smoothing.py:62-127          label-smoothed CE: nll == plain CE, padding ignored, per-row reduction,
                                                zero smoothing == CE
(the beam-search tables of tests/test_sequence_generator.py are in tests/test_beam_search.py)."""
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
