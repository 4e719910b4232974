"""GPU end-to-end parity: B200 encoder model + CTC criterion + trainer vs fixtures recorded from the REAL
reference model (tests/golden/encoder_*.npz) and vs the oracle on fresh inputs.  Tolerances: bf16 compute vs
the fp32 reference -- logits within 6% of the logit range, loss within 3%, gradients within 12% of each
tensor's max (medians far lower); dropout = 0 for value parity (custom RNG streams cannot match ATen's),
dropout > 0 covered by mask-consistency tests."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Dict:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def pad(self):
        return 1

    def eos(self):
        return 2

    def index(self, sym):
        return 0


class _Task:
    feat_dim, feat_in_channels = 80, 1

    def __init__(self, V):
        self.target_dictionary = _Dict(V)


def _build(layer_type, g, dropout=0.0, dev="cuda:0"):
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel

    cfg = SpeechTransformerConfig.from_dict(dict(
        dropout=dropout, attention_dropout=dropout, activation_dropout=dropout, layernorm_embedding=True,
        max_source_positions=3600,
        encoder=dict(embed_dim=64, ffn_embed_dim=128, layers=2, attention_heads=4, normalize_before=True,
                     learned_pos="learned" in layer_type,
                     share_learned_relative_positional_embeddings_across_heads=layer_type.endswith("_sh"),
                     relative_positional_embeddings=True, layer_type=layer_type.split("_")[0],
                     depthwise_conv_kernel_size=31)))
    m = SpeechTransformerEncoderModel.build_model(cfg, _Task(50))
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    return m.finalize_(torch.device(dev))


def _sample(g, dev):
    return {"net_input": {"src_tokens": torch.from_numpy(g["feats"]).to(dev), "src_lengths": torch.from_numpy(g["lens"]).to(dev)},
            "target": torch.from_numpy(g["target"]).to(dev), "ntokens": 13}


@pytest.mark.parametrize("layer_type", ["conformer", "transformer", "transformer_learned", "conformer_learned_sh"])
def test_encoder_vs_reference_fixture(layer_type, golden_dir, parity):
    from espresso_b200 import lib
    from espresso_b200.criterions import CtcLossCriterion

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "encoder_%s.npz" % layer_type))
    m = _build(layer_type, g)
    crit = CtcLossCriterion(_Task(50))
    sample = _sample(g, dev)
    n0 = lib.launch_count()
    m.train()
    m.flat.zero_grad()
    loss, sample_size, log = crit(m, sample)
    loss.backward()
    m.encoder.sync_torch_grads_()
    torch.cuda.synchronize()
    assert lib.launch_count() - n0 > 50  # the native kernels really ran
    parity("encoder %s: CTC loss rel err vs reference fp32" % layer_type, abs(loss.item() - float(g["loss_train"])) / float(g["loss_train"]), 5e-3)
    worst = []
    for k in g.files:
        if not k.startswith("grad.encoder."):
            continue
        name = k[len("grad."):]
        if ("pre_encoder.convolutions" in name and name.endswith(".bias")) or name.endswith("k_proj.bias"):
            continue  # analytically zero gradients
        ours = m.flat.grad(name).cpu().numpy()
        refg = g[k]
        worst.append((np.linalg.norm(ours - refg) / max(np.linalg.norm(refg), 1e-3), name))
    worst.sort(reverse=True)
    parity("encoder %s: worst gradient rel-Frobenius (%s)" % (layer_type, worst[0][1]), worst[0][0], 0.25)
    parity("encoder %s: worst gradient outside the conv front" % layer_type, max(w[0] for w in worst if "pre_encoder" not in w[1]), 0.1)
    parity("encoder %s: median gradient rel-Frobenius" % layer_type, np.median([w[0] for w in worst]), 0.03)
    m.eval()
    with torch.no_grad():
        net = m(**sample["net_input"])
    logits = net["encoder_out"][0].transpose(0, 1).float().cpu().numpy()
    ref = g["logits_eval"]
    assert logits.shape == ref.shape and np.array_equal(net["src_lengths"][0].cpu().numpy(), g["out_lens"])
    parity("encoder %s: eval logits max abs / max |ref|" % layer_type, np.abs(logits - ref).max() / np.abs(ref).max(), 2e-2)
    parity("encoder %s: eval logits rel-Frobenius" % layer_type, np.linalg.norm(logits - ref) / np.linalg.norm(ref), 1.5e-2)


def test_training_reduces_loss_with_dropout(golden_dir):
    """A few updates with dropout 0.1 through the B200 Trainer on the fixture batch: loss must go down and stay
    finite (exercises every dropout path forward+backward with regenerated masks)."""
    from espresso_b200.criterions import CtcLossCriterion
    from espresso_b200.optim import NoamLRScheduler
    from espresso_b200.trainer import Trainer

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "encoder_conformer.npz"))
    m = _build("conformer", g, dropout=0.1)
    tr = Trainer(m, CtcLossCriterion(_Task(50)), NoamLRScheduler(0.5, 5, 64, 1e-6), clip_norm=2.0)
    sample = _sample(g, dev)
    losses = []
    for _ in range(25):
        tr.train_step([sample])
        losses.append(tr.stats()["loss"])
    assert all(np.isfinite(losses)), losses
    assert min(losses[-5:]) < 0.9 * losses[0], losses
    assert tr.stats()["sample_size"] == 3


def test_graph_replay_follows_lengths_and_buckets(golden_dir):
    """One captured CUDA graph must serve every batch of its shape bucket: batches of the same padded shape but
    different length patterns (incl. one whose subsampled lengths are all equal, i.e. no padding after the conv
    front) and batches whose raw shape is smaller than the bucket.  Gradients of the graphed step == eager step."""
    from espresso_b200.criterions import CtcLossCriterion
    from espresso_b200.optim import NoamLRScheduler
    from espresso_b200.trainer import Trainer

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "encoder_conformer.npz"))

    def batch(T, lens, U=6, seed=0):
        rs = np.random.RandomState(seed)
        B = len(lens)
        feats = rs.randn(B, T, 80).astype(np.float32)
        for b, l in enumerate(lens):
            feats[b, l:] = 0
        tgt = np.full((B, U + 1), 1, dtype=np.int64)
        for b in range(B):
            u = 2 + (b + seed) % (U - 1)
            tgt[b, :u] = rs.randint(4, 50, size=u)
            tgt[b, u] = 2
        lt = torch.tensor(lens)
        return {"net_input": {"src_tokens": torch.from_numpy(feats).to(dev), "src_lengths": lt.to(dev), "src_lengths_cpu": lt},
                "target": torch.from_numpy(tgt).to(dev)}

    m = _build("conformer", g)
    crit = CtcLossCriterion(_Task(50))
    eager = Trainer(m, crit, NoamLRScheduler(0.0, 5, 64, 0.0), use_cuda_graphs=False)
    graphed = Trainer(m, crit, NoamLRScheduler(0.0, 5, 64, 0.0), use_cuda_graphs=True, bucket_frames=32, bucket_tokens=8)
    graphed._reduce_and_update = lambda: None  # keep the parameters fixed: compare raw gradient buffers
    # bucket = 128 frames.  [128,127,126]: subsampled lengths all 32 (no pads after the conv front); others padded.
    cases = [batch(128, [128, 127, 126], seed=1), batch(128, [128, 127, 126], seed=2), batch(128, [128, 90, 41], seed=3),
             batch(128, [128, 127, 126], seed=4), batch(117, [117, 60, 33], U=5, seed=5), batch(100, [100, 99, 98], seed=6)]
    for i, smp in enumerate(cases):
        graphed.train_step([smp])
        torch.cuda.synchronize()
        got = m.flat.g32.clone()
        # eager reference on the batch padded by hand to the bucket shape (zero frames, pad tokens): the bucketed batch
        # carries extra all-zero frames that BatchNorm batch statistics see, exactly as the reference would if its
        # collater padded to the same multiple
        x, t = smp["net_input"]["src_tokens"], smp["target"]
        xp = torch.zeros(x.shape[0], 128, 80, device=dev)
        xp[:, : x.shape[1]] = x
        tp = torch.full((t.shape[0], 8), 1, dtype=t.dtype, device=dev)
        tp[:, : t.shape[1]] = t
        eager._fwd_bwd([{"net_input": dict(smp["net_input"], src_tokens=xp), "target": tp}])
        torch.cuda.synchronize()
        ref = m.flat.g32.clone()
        denom = ref[: m.flat.numel].abs().max().item()
        err = (got - ref)[: m.flat.numel].abs().max().item() / denom
        assert err < 2e-2, (i, err)
        assert abs(got[m.flat.numel + 3].item() - ref[m.flat.numel + 3].item()) < 2e-2 * abs(ref[m.flat.numel + 3].item())
    assert graphed.graph_hits >= 2 and len(graphed._graphs) <= 3, (graphed.graph_hits, len(graphed._graphs))


def test_full_size_layer_shapes():
    """cfg-3 shapes (d=512, ffn=2048, H=8, k=31, V=5004) on a LibriSpeech-shaped batch: forward/backward run,
    outputs are finite, padded rows of the input stay inert for the logits of other utterances."""
    from espresso_b200.criterions import CtcLossCriterion
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = SpeechTransformerConfig.from_dict(dict(
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, layernorm_embedding=True,
        encoder=dict(embed_dim=512, ffn_embed_dim=2048, layers=2, attention_heads=8, normalize_before=True, learned_pos=False,
                     relative_positional_embeddings=True, layer_type="conformer", depthwise_conv_kernel_size=31)))
    m = SpeechTransformerEncoderModel.build_model(cfg, _Task(5004)).finalize_(dev)
    B, T = 6, 1203
    lens = torch.tensor([1203, 1100, 900, 640, 333, 100], device=dev)
    feats = torch.randn(B, T, 80, device=dev)
    for b in range(B):
        feats[b, lens[b]:] = 0
    tgt = torch.full((B, 30), 1, dtype=torch.long, device=dev)
    for b in range(B):
        u = 5 + 3 * b
        tgt[b, :u] = torch.randint(4, 5004, (u,), device=dev)
        tgt[b, u] = 2
    crit = CtcLossCriterion(_Task(5004))
    m.train()
    m.flat.zero_grad()
    loss, ss, log = crit(m, {"net_input": {"src_tokens": feats, "src_lengths": lens}, "target": tgt})
    loss.backward()
    m.encoder.sync_torch_grads_()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item()
    assert torch.isfinite(m.flat.grads).all().item()
    assert m.flat.grads.abs().max().item() > 0


def test_encdec_vs_reference_fixture(golden_dir):
    """speech_transformer_base (Transformer encoder + decoder) + label-smoothed CE on the GPU vs the fixture recorded
    from the real reference model."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_orchestration import _build_encdec
    from espresso_b200.criterions import LabelSmoothedCrossEntropyV2Criterion

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "encdec_transformer.npz"))
    m = _build_encdec(g).finalize_(dev)
    crit = LabelSmoothedCrossEntropyV2Criterion(_Task(50), label_smoothing=float(g["eps"]))
    sample = {"net_input": {"src_tokens": torch.from_numpy(g["feats"]).to(dev), "src_lengths": torch.from_numpy(g["lens"]).to(dev),
                            "prev_output_tokens": torch.from_numpy(g["prev_output_tokens"]).to(dev)},
              "target": torch.from_numpy(g["target"]).to(dev)}
    m.train()
    m.flat.zero_grad()
    loss, sample_size, log = crit(m, sample)
    loss.backward()
    m.encoder.sync_torch_grads_()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) < 0.03 * float(g["loss"])
    assert abs(log["nll_loss"].item() - float(g["nll"])) < 0.03 * float(g["nll"])
    worst = []
    for k in g.files:
        if not k.startswith("grad."):
            continue
        name = k[len("grad."):]
        if ("pre_encoder.convolutions" in name and name.endswith(".bias")) or name.endswith("k_proj.bias"):
            continue
        ours, refg = m.flat.grad(name).cpu().numpy(), g[k]
        worst.append((np.linalg.norm(ours - refg) / max(np.linalg.norm(refg), 1e-3), name))
    worst.sort(reverse=True)
    assert worst[0][0] < 0.25, worst[:5]
    assert max(w[0] for w in worst if "pre_encoder" not in w[1]) < 0.1, worst[:8]


def test_lsce_embed_argmax_kernels():
    from espresso_b200 import ops
    from oracle import ops_ref as O

    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    for V in (50, 5004, 9000):
        ld = (V + 7) // 8 * 8
        R = 37
        x = (torch.randn(R, ld) * 2).bfloat16()
        t = torch.randint(0, V, (R,), dtype=torch.int32)
        t[::5] = 1  # pad rows
        loss, nll, grad = ops.lsce_loss(x.to(dev), V, t.to(dev), 1, 0.1, grad_scale=0.5)
        lr, nr, gr = O.lsce_loss(x, V, t, 1, 0.1, grad_scale=0.5)
        assert torch.allclose(loss.cpu(), lr, rtol=1e-4, atol=1e-3) and torch.allclose(nll.cpu(), nr, rtol=1e-4, atol=1e-3)
        assert (grad.float().cpu() - gr.float()).abs().max().item() < 4e-3
        assert not grad[:, V:].any()
        am = ops.argmax_rows(x.to(dev), V)
        assert torch.equal(am.cpu(), O.argmax_rows(x, V))
        # unigram and temporal smoothing (rows viewed as [B, U] for the temporal neighbours)
        uni = torch.rand(V) + 0.01
        uni = uni / uni.sum()
        for mode, Uu in ((ops.SMOOTH_UNIGRAM, 0), (ops.SMOOTH_TEMPORAL, 37), (ops.SMOOTH_TEMPORAL, 1)):
            kw = dict(smoothing=mode, unigram=uni.to(dev) if mode == ops.SMOOTH_UNIGRAM else None, U=Uu)
            kwr = dict(smoothing=mode, unigram=uni, U=Uu)
            loss, nll, grad = ops.lsce_loss(x.to(dev), V, t.to(dev), 1, 0.1, grad_scale=0.5, **kw)
            lr, nr, gr = O.lsce_loss(x, V, t, 1, 0.1, grad_scale=0.5, **kwr)
            assert torch.allclose(loss.cpu(), lr, rtol=1e-4, atol=2e-3), (V, mode)
            assert torch.allclose(nll.cpu(), nr, rtol=1e-4, atol=1e-3)
            assert (grad.float().cpu() - gr.float()).abs().max().item() < 4e-3, (V, mode)
    # the reference's own outputs for the three smoothing types (tests/golden/label_smoothing.npz)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "label_smoothing.npz"))
    lg = torch.from_numpy(g["logits"])
    Bq, Uq, Vq = lg.shape
    xq = torch.zeros(Bq * Uq, (Vq + 7) // 8 * 8, dtype=torch.bfloat16)
    xq[:, :Vq] = lg.view(-1, Vq).bfloat16()
    tq = torch.from_numpy(g["target"]).view(-1).int()
    for name, mode in (("uniform", 0), ("unigram", 1), ("temporal", 2)):
        loss, nll, grad = ops.lsce_loss(xq.to(dev), Vq, tq.to(dev), int(g["pad"]), float(g["eps"]), smoothing=mode,
                                        unigram=torch.from_numpy(g["unigram"]).to(dev), U=Uq)
        assert abs(loss.sum().item() - float(g["loss_" + name])) < 2e-4 * abs(float(g["loss_" + name])), name
        assert abs(nll.sum().item() - float(g["nll_" + name])) < 2e-4 * abs(float(g["nll_" + name])), name
        assert np.abs(grad[:, :Vq].float().cpu().view(Bq, Uq, Vq).numpy() - g["grad_" + name]).max() < 4e-3, name
    E = torch.randn(50, 64).bfloat16()
    pos = torch.randn(9, 64).bfloat16()
    tok = torch.randint(0, 50, (4 * 9,), dtype=torch.int32)
    tok[3] = 1
    y = ops.embed_fwd(tok.to(dev), E.to(dev), pos.to(dev), 9, 8.0, 1)
    assert (y.float().cpu() - O.embed_fwd(tok, E, pos, 9, 8.0, 1).float()).abs().max().item() < 0.07
    dx = torch.randn(36, 64).bfloat16()
    dE, dEr = torch.zeros(50, 64, device=dev), torch.zeros(50, 64)
    ops.embed_bwd(tok.to(dev), dx.to(dev), dE, 8.0, 1)
    O.embed_bwd(tok, dx, dEr, 8.0, 1)
    assert (dE.cpu() - dEr).abs().max().item() < 1e-3
    # rectangular + causal softmax
    H, B, Tq, Tk = 2, 3, 9, 21
    ld = 24
    s = (torch.randn(H, B, Tq, ld) * 2).bfloat16()
    lens = torch.tensor([21, 13, 5], dtype=torch.int32)
    p, _ = ops.attn_softmax_fwd(s.to(dev), Tk, lens.to(dev))
    pr, _ = O.attn_softmax_fwd(s, Tk, lens)
    assert (p.float().cpu() - pr.float()).abs().max().item() < 0.01
    sq = (torch.randn(H, B, 16, 16) * 2).bfloat16()
    l2 = torch.tensor([16, 9, 3], dtype=torch.int32)
    p, _ = ops.attn_softmax_fwd(sq.to(dev), 16, l2.to(dev), causal=True)
    pr, _ = O.attn_softmax_fwd(sq, 16, l2, causal=True)
    assert (p.float().cpu() - pr.float()).abs().max().item() < 0.01
    dp = torch.randn(H, B, Tq, ld).bfloat16()
    ds, _ = ops.attn_softmax_bwd(ops.attn_softmax_fwd(s.to(dev), Tk, lens.to(dev))[0], dp.to(dev), Tk, 0, want_dbd=False)
    dsr, _ = O.attn_softmax_bwd(O.attn_softmax_fwd(s, Tk, lens)[0], dp, Tk, 0, want_dbd=False)
    assert (ds.float().cpu()[..., :Tk] - dsr.float()[..., :Tk]).abs().max().item() < 0.03


def test_transducer_vs_reference_fixture(golden_dir):
    """Conformer-Transducer (encoder engine + cuDNN LSTM predictor + native joint / fc_out GEMM / RNN-T loss) on the GPU
    vs the fixture recorded from the real reference model + torchaudio rnnt_loss."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_orchestration import _build_transducer
    from espresso_b200.criterions import TransducerLossCriterion

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "transducer_conformer.npz"))
    m = _build_transducer(g).finalize_(dev)
    crit = TransducerLossCriterion(_Task(50))
    sample = {"net_input": {"src_tokens": torch.from_numpy(g["feats"]).to(dev), "src_lengths": torch.from_numpy(g["lens"]).to(dev),
                            "prev_output_tokens": torch.from_numpy(g["prev_output_tokens"]).to(dev)},
              "target": torch.from_numpy(g["target"]).to(dev)}
    m.train()
    m.flat.zero_grad()
    loss, sample_size, log = crit(m, sample)
    loss.backward()
    m.sync_torch_grads_()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) < 0.03 * float(g["loss"])
    worst = []
    for k in g.files:
        if not k.startswith("grad."):
            continue
        name = k[len("grad."):]
        if ("pre_encoder.convolutions" in name and name.endswith(".bias")) or name.endswith("k_proj.bias"):
            continue
        ours, refg = m.flat.grad(name).cpu().numpy(), g[k]
        worst.append((np.linalg.norm(ours - refg) / max(np.linalg.norm(refg), 1e-3), name))
    worst.sort(reverse=True)
    assert worst[0][0] < 0.25, worst[:5]
    assert max(w[0] for w in worst if "pre_encoder" not in w[1]) < 0.12, worst[:8]


def test_full_size_encoder_value_parity(golden_dir):
    """VALUE parity at the benchmarked configuration (17 x 512 Conformer, V = 5004, 3.1 / 6.4 / 10 s utterances) on the CUDA
    path: logits, log-normalisers, CTC loss and a spread of gradients are as close to the fp32 reference as the
    reference's own bf16 run (tests/fullsize_util.py states the contract; achieved errors are printed)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fullsize_util import run_and_check

    run_and_check("cuda:0", golden_dir)


def test_chunk_streaming_encoder_vs_reference_fixture(golden_dir):
    """Chunk-streaming Conformer (chunk_size / chunk_left_window / chunk_right_window) against the REAL reference model
    (tests/golden/encoder_streaming.npz, oracle/pin_against_reference.py::pin_streaming): training at num_updates 0
    and 1 (the reference draws the first-or-last partial chunk under numpy_seed(num_updates); the fixture holds both
    outcomes) and eval (always partial in last).  The masks run inside the fused attention kernel as key ranges."""
    from espresso_b200.criterions import CtcLossCriterion
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "encoder_streaming.npz"))
    chunk, lw, rw = (int(c) for c in g["chunk"])
    cfg = SpeechTransformerConfig.from_dict(dict(
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, layernorm_embedding=True, max_source_positions=3600,
        encoder=dict(embed_dim=128, ffn_embed_dim=128, layers=2, attention_heads=2, normalize_before=True,
                     relative_positional_embeddings=True, layer_type="conformer", depthwise_conv_kernel_size=31,
                     conv_channels="[16, 16, 32, 32]", chunk_size=chunk, chunk_left_window=lw, chunk_right_window=rw)))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    crit = CtcLossCriterion(_Task(50))
    sample = _sample(g, dev)
    for mode, nu in (("train", 0), ("train", 1), ("eval", 7)):
        tag = "%s%d" % (mode, nu)
        m = SpeechTransformerEncoderModel.build_model(cfg, _Task(50))
        m.load_state_dict(sd, strict=True)
        m.finalize_(dev)
        m.train(mode == "train")
        m.set_num_updates(nu)
        assert m.encoder.has_attn_mask
        if mode == "train":
            m.flat.zero_grad()
            loss, _, _ = crit(m, sample)
            loss.backward()
            m.encoder.sync_torch_grads_()
        else:
            with torch.no_grad():
                loss, _, _ = crit(m, sample)
        ref_loss = float(g["loss_" + tag])
        assert abs(loss.item() - ref_loss) < 0.03 * ref_loss, (tag, loss.item(), ref_loss)
        # the two training coins give different losses in the fixture; ours must follow the right one
        lo, hi = m.encoder.attn_key_bounds(int(g["out_lens"].max()), int(g["out_lens"].max()))
        j = np.arange(len(lo))[None, :]
        assert np.array_equal((j < lo[:, None]) | (j >= hi[:, None]), g["hidden_" + tag]), tag
        if mode == "train" and nu == 0:
            worst = []
            for k in g.files:
                if not k.startswith("grad_train0.encoder."):
                    continue
                name = k[len("grad_train0."):]
                if ("pre_encoder.convolutions" in name and name.endswith(".bias")) or name.endswith("k_proj.bias"):
                    continue
                ours = m.flat.grad(name).cpu().numpy()
                worst.append((np.linalg.norm(ours - g[k]) / max(np.linalg.norm(g[k]), 1e-3), name))
            worst.sort(reverse=True)
            assert worst[0][0] < 0.25, worst[:5]
            assert max(w[0] for w in worst if "pre_encoder" not in w[1]) < 0.1, worst[:8]
            assert np.median([w[0] for w in worst]) < 0.03
        if mode == "eval":
            with torch.no_grad():
                net = m(**sample["net_input"])
            logits = net["encoder_out"][0].transpose(0, 1).float().cpu().numpy()
            ref = g["logits_" + tag]
            assert np.abs(logits - ref).max() < 0.06 * np.abs(ref).max()
