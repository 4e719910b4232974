"""Shared by the CPU (host orchestration over oracle/ops_ref) and GPU (CUDA kernels) parity tests at the BENCHMARKED
configuration: 17 x 512 Conformer + CTC, V = 5004, against tests/golden/fullsize_conformer.npz, which holds outputs of the
REAL reference model in fp32 (truth) and in bf16 (`model.bfloat16()`: the reference's own bf16 error = the yardstick).

Contract asserted (and printed): our bf16 path is as close to the fp32 reference as the reference's bf16 path is --
err_ours <= 1.25 x err_reference_bf16 (+ a small floor) on logits, log-normalisers and every stored gradient, the loss
within 2e-3 relative (the reference's own bf16 loss is 5.5e-4 away), and all within SURVEY 8d's 2e-2 ... 5e-2 vs fp32."""
import os

import numpy as np
import torch

from oracle import conformer as OC
from oracle.fullsize import FULLSIZE_CFG, FULLSIZE_GRADS, FULLSIZE_GRADS_SUB, fullsize_cotangent, fullsize_inputs


class _Dict:
    def __len__(self):
        return 5004

    def pad(self):
        return 1

    def eos(self):
        return 2

    def index(self, sym):
        return 0


class _Task:
    feat_dim, feat_in_channels, target_dictionary = 80, 1, _Dict()


def build_model(device):
    from espresso_b200.models import SpeechTransformerConfig, SpeechTransformerEncoderModel

    cfg = SpeechTransformerConfig.from_dict(dict(
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, layernorm_embedding=True, max_source_positions=3600,
        encoder=dict(embed_dim=512, ffn_embed_dim=2048, layers=17, attention_heads=8, normalize_before=True, learned_pos=False,
                     relative_positional_embeddings=True, layer_type="conformer", depthwise_conv_kernel_size=31)))
    m = SpeechTransformerEncoderModel.build_model(cfg, _Task())
    sd = OC.random_state_dict(FULLSIZE_CFG, seed=1)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("num_batches_tracked" in k or k.endswith("version") or k.endswith("_float_tensor") for k in res.missing_keys), res.missing_keys
    return m.finalize_(torch.device(device))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run_and_check(device, golden_dir, slack=1.25):
    from espresso_b200.criterions import CtcLossCriterion

    g = np.load(os.path.join(golden_dir, "fullsize_conformer.npz"))
    feats, lens, tgt = fullsize_inputs()
    dev = torch.device(device)
    m = build_model(dev)
    crit = CtcLossCriterion(_Task(), zero_infinity=True, sentence_avg=True)
    sample = {"net_input": {"src_tokens": torch.from_numpy(feats).to(dev), "src_lengths": torch.from_numpy(lens).to(dev),
                            "src_lengths_cpu": torch.from_numpy(lens)},
              "target": torch.from_numpy(tgt).to(dev)}
    m.train()
    # (1) CTC loss value through the criterion (forward only matters here)
    m.flat.zero_grad()
    loss, sample_size, log = crit(m, sample)
    loss = loss.detach()
    # (2) gradients of the well-conditioned linear functional sum(G * logits) (oracle/fullsize.py): one more training
    # forward, then the hand-written backward driven by dlogits = G
    m.flat.zero_grad()
    for mod in m.modules():  # the fixture's BatchNorm running statistics are not compared; keep the run repeatable
        if hasattr(mod, "num_batches_tracked"):
            pass
    net = m(**sample["net_input"])
    out_full = net["b200_out"]                                  # [B, T', ldV] batch-major logits
    out = net["encoder_out"][0].transpose(0, 1).detach()        # [B, T', V]
    G = torch.from_numpy(fullsize_cotangent(g["out_lens"].tolist()))
    Gp = torch.zeros(out_full.shape, dtype=out_full.dtype, device=dev)
    Gp[..., : G.shape[-1]] = G.to(dev)
    out_full.backward(Gp)
    m.encoder.sync_torch_grads_()
    # (3) the CTC kernel's gradient against autograd of the reference's call (ctc_loss.py:85-94) on IDENTICAL logits
    ctc_grad_check(m, crit, sample, out, dev)
    report = []

    def check(name, ours, ref32, ref16, floor, cap=8e-2):
        e_ours, e_ref = rel(ours, ref32), rel(ref16, ref32)
        report.append((name, e_ours, e_ref))
        assert e_ours <= slack * e_ref + floor, "%s: ours %.3g vs reference-bf16 %.3g (both against fp32)" % (name, e_ours, e_ref)
        assert e_ours <= cap, (name, e_ours)

    lg = out.float().cpu().numpy()
    assert np.array_equal(g["out_lens"], m.output_lengths(torch.from_numpy(lens)).numpy())
    check("logits", lg[:, ::5, ::11], g["logits_sub_fp32"], g["logits_sub_bf16"], 5e-3)
    lse = torch.logsumexp(out.float(), dim=-1).cpu().numpy()
    check("log-normaliser", lse, g["lse_fp32"], g["lse_bf16"], 2e-3)
    l_ours, l32, l16 = float(loss.item()), float(g["loss_fp32"]), float(g["loss_bf16"])
    report.append(("loss", abs(l_ours - l32) / l32, abs(l16 - l32) / l32))
    assert abs(l_ours - l32) <= 2e-3 * l32, (l_ours, l32, l16)
    for n in FULLSIZE_GRADS:
        check("grad " + n, m.flat.grad(n).float().cpu().numpy(), g["grad_fp32." + n], g["grad_bf16." + n], 2e-2, cap=0.25)
    for n in FULLSIZE_GRADS_SUB:
        gr = m.flat.grad(n).float().cpu()
        gr = gr.reshape(gr.shape[0], -1)[::8, ::8].numpy()
        check("grad[::8,::8] " + n, gr, g["gradsub_fp32." + n], g["gradsub_bf16." + n], 2e-2, cap=0.25)
    gn = float(torch.sqrt((m.flat.grads.double() ** 2).sum()).item())
    report.append(("|grad| (all parameters)", abs(gn - float(g["gnorm_fp32"])) / float(g["gnorm_fp32"]),
                   abs(float(g["gnorm_bf16"]) - float(g["gnorm_fp32"])) / float(g["gnorm_fp32"])))
    # measured: reference bf16 -0.9 %, host path over oracle/ops_ref -1.1 %, CUDA path -0.7 % (round-1 attention chain) /
    # -2.0 % (fused attention): every bf16 run lands below the fp32 norm; per-tensor errors above are the real check
    assert abs(gn - float(g["gnorm_fp32"])) <= 3e-2 * float(g["gnorm_fp32"]), (gn, float(g["gnorm_fp32"]))
    print("\nfull-size parity (relative error against the fp32 reference):   ours    | reference bf16")
    for name, a, b in report:
        print("  %-62s %.3e | %.3e" % (name, a, b))
    return report


def ctc_grad_check(m, crit, sample, logits, dev):
    import torch.nn.functional as Fn

    from espresso_b200 import ops
    from espresso_b200.criterions.ctc_loss import compact_targets

    tgt = sample["target"]
    B, T, V = logits.shape
    olens = m.output_lengths(sample["net_input"]["src_lengths_cpu"])
    lg = logits.float().cpu().clone().requires_grad_(True)
    keep = (tgt.cpu() != 1) & (tgt.cpu() != 2)
    with torch.backends.cudnn.flags(enabled=False):
        ref = Fn.ctc_loss(Fn.log_softmax(lg, -1).transpose(0, 1), tgt.cpu().masked_select(keep), olens, keep.sum(-1), blank=0,
                          reduction="sum", zero_infinity=True)
    ref.backward()
    tg, tl = compact_targets(tgt, 1, 2)
    ld = (V + 7) // 8 * 8
    lb = torch.zeros(B, T, ld, dtype=torch.bfloat16, device=dev)
    lb[..., :V] = logits.to(torch.bfloat16)
    loss_b, grad = ops.ctc_loss(lb, V, olens.to(torch.int32).to(dev), tg, tl, 0)
    e_loss = abs(float(loss_b.sum()) - float(ref.detach())) / float(ref.detach())
    e_grad = rel(grad.float().cpu().numpy()[..., :V], lg.grad.numpy())
    print("\nCTC at V=%d, T'=%d on identical logits: loss rel %.2e, gradient rel-Frobenius %.2e (bf16 storage)" % (V, T, e_loss, e_grad))
    assert e_loss < 1e-5 and e_grad < 4e-3
