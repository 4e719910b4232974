"""Chunk-streaming / limited-context self-attention masks on the host (no GPU): the per-row key ranges the fused
attention kernel consumes, against masks recorded from the REAL reference (tests/golden/encoder_streaming.npz, written by
oracle/pin_against_reference.py::pin_streaming, which also checks 1138 masks live against
espresso/tools/utils.py:131-194) and against the oracle's two statements of masked attention."""
import os

import numpy as np
import pytest
import torch


class _Dict:
    def __len__(self):
        return 50

    def pad(self):
        return 1


class _Task:
    feat_dim, feat_in_channels, target_dictionary = 80, 1, _Dict()


def _model(**enc):
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel

    cfg = SpeechTransformerConfig.from_dict(dict(
        layernorm_embedding=True, max_source_positions=3600,
        encoder=dict(embed_dim=128, ffn_embed_dim=128, layers=1, attention_heads=2, normalize_before=True,
                     relative_positional_embeddings=True, layer_type="conformer", conv_channels="[16, 16, 32, 32]", **enc)))
    return SpeechTransformerEncoderModel.build_model(cfg, _Task())


def test_model_key_ranges_reproduce_reference_masks(golden_dir):
    from espresso_b200.tools.utils import bounds_to_mask

    g = np.load(os.path.join(golden_dir, "encoder_streaming.npz"))
    chunk, lw, rw = (int(c) for c in g["chunk"])
    m = _model(chunk_size=chunk, chunk_left_window=lw, chunk_right_window=rw)
    T = int(g["out_lens"].max())
    seen = set()
    for mode, nu in (("train", 0), ("train", 1), ("eval", 7)):
        m.train(mode == "train")
        m.set_num_updates(nu)
        state = np.random.get_state()[1][:4].copy()
        lo, hi = m.encoder.attn_key_bounds(T, T)
        assert np.array_equal(np.random.get_state()[1][:4], state)   # numpy_seed restores the global RNG
        assert lo.dtype == np.int32 and hi.dtype == np.int32
        assert np.array_equal(~bounds_to_mask(lo, hi), g["hidden_%s%d" % (mode, nu)])
        seen.add((mode, int(hi[0])))
    assert len(seen) == 3 or len({s for s in seen if s[0] == "train"}) == 2   # both partial-chunk placements occurred
    # bucketed shapes (T larger than the longest utterance): extra rows are padding with a non-empty range
    lo2, hi2 = m.encoder.attn_key_bounds(T, T + 23)
    assert len(lo2) == T + 23 and np.array_equal(lo2[:T], lo) and (hi2[T:] > lo2[T:]).all() and hi2.max() <= T


@pytest.mark.parametrize("ctx,expect_lo,expect_hi", [
    ("(2, 1)", [0, 0, 0, 1, 2], [2, 3, 4, 5, 5]),
    ("(None, 0)", [0, 0, 0, 0, 0], [1, 2, 3, 4, 5]),
    ("(0, None)", [0, 1, 2, 3, 4], [5, 5, 5, 5, 5]),
])
def test_transformer_context_band(ctx, expect_lo, expect_hi):
    m = _model(transformer_context=ctx)
    assert m.encoder.has_attn_mask
    lo, hi = m.encoder.attn_key_bounds(5, 5)
    assert lo.tolist() == expect_lo and hi.tolist() == expect_hi


def test_no_mask_configurations():
    assert not _model().encoder.has_attn_mask
    assert not _model(transformer_context="(None, None)").encoder.has_attn_mask
    assert _model().encoder.attn_key_bounds(9, 9) is None
    with pytest.raises(ValueError):
        _model(transformer_context="(1, -2)")


def test_oracle_statements_of_masked_attention_agree():
    """oracle/ops_ref.attn_fused_fwd (the kernel's checker, key ranges) and oracle/conformer.relpos_mha (pinned against the
    reference model with attn_mask) describe the same function."""
    from espresso_b200.tools.utils import bounds_to_mask, chunk_streaming_bounds
    from oracle import conformer as O
    from oracle import ops_ref

    torch.manual_seed(3)
    B, T, H, d = 2, 37, 2, 128
    np.random.seed(0)
    lo, hi = chunk_streaming_bounds(T, 5, 1, 1)
    hidden = torch.from_numpy(~bounds_to_mask(lo, hi))
    sd = {"q_proj.weight": torch.randn(d, d) * 0.1, "k_proj.weight": torch.randn(d, d) * 0.1, "v_proj.weight": torch.randn(d, d) * 0.1,
          "out_proj.weight": torch.eye(d), "q_proj.bias": torch.zeros(d), "k_proj.bias": torch.zeros(d),
          "v_proj.bias": torch.zeros(d), "out_proj.bias": torch.zeros(d), "pos_bias_u": torch.randn(d) * 0.1,
          "pos_bias_v": torch.randn(d) * 0.1, "pos_proj.weight": torch.randn(d, d) * 0.1}
    x = torch.randn(B, T, d)
    lens = torch.tensor([T, 22])
    pad = torch.arange(T)[None, :] >= lens[:, None]
    add = hidden.float().masked_fill(hidden, -1e8)
    want = O.relpos_mha(sd, "", x, pad, H, 0.0, False, add)
    s = (d // H) ** -0.5
    q = x @ sd["q_proj.weight"].t()
    qu = ((q + sd["pos_bias_u"]) * s).reshape(B * T, d)
    qv = ((q + sd["pos_bias_v"]) * s).reshape(B * T, d)
    k = (x @ sd["k_proj.weight"].t()).reshape(B * T, d)
    v = (x @ sd["v_proj.weight"].t()).reshape(B * T, d)
    pos = O.rel_pos_table(T, d, x.dtype) @ sd["pos_proj.weight"].t()
    got, _, _ = ops_ref.attn_fused_fwd(qu, qv, k, v, pos, B, T, H, lens, key_bounds=(torch.from_numpy(lo), torch.from_numpy(hi)))
    rows = (~pad)  # padded query rows whose visible chunk is all padding are degenerate (compared on the GPU side only)
    err = (got.float().view(B, T, d) - want)[rows].abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item()), err
