"""Pin the oracle against the REAL reference and (re)generate tests/golden/*.npz.

Runs only in the authoring container (needs /root/reference; imported through oracle/refshim).
    python -m oracle.pin_against_reference [frontend] [ctc] [conformer] ...
Each section (1) runs the reference's own code on seeded synthetic inputs, (2) asserts the oracle
restatement agrees, (3) stores inputs + reference outputs as a small fixture.  The GPU tests compare
the CUDA path against those fixtures (and against the oracle on fresh seeded inputs).
"""
import os
import sys

import numpy as np
import torch

from oracle import refshim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SPECAUG_CFG = {"time_warp_W": 0, "freq_mask_F": 27, "freq_mask_N": 2, "time_mask_pm": 0.04, "time_mask_ps": 0.04}


def pin_frontend():
    from espresso.data.feature_transforms.adaptive_specaugment import AdaptiveSpecAugmentTransform
    from espresso.tools.utils import get_torchaudio_fbank_or_mfcc
    from fairseq.data import data_utils
    from fairseq.data.audio.feature_transforms.global_cmvn import GlobalCMVN  # noqa: F401

    from oracle import frontend as O

    durs = [1.0, 2.35, 0.5, 3.17, 0.03]
    waves = [O.synth_waveform(i, d) for i, d in enumerate(durs)]
    # CMVN stats with the reference formula (espresso/tools/compute_global_cmvn_stats.py:94-118)
    feats_ref = [get_torchaudio_fbank_or_mfcc(w[None, :], 16000, n_bins=80) for w in waves[:4]]
    allf = np.concatenate(feats_ref, axis=0).astype(np.float64)
    mean = allf.mean(axis=0)
    var = (allf ** 2).mean(axis=0) - mean ** 2
    std = np.sqrt(np.maximum(var, 1e-8))
    aug = AdaptiveSpecAugmentTransform.from_config_dict(SPECAUG_CFG)
    out = {"durs": np.array(durs), "cmvn_mean": mean, "cmvn_std": std}
    worst = 0.0
    for i, w in enumerate(waves):
        n = len(w)
        if O.num_frames(n) == 0:
            continue
        ref_fb = get_torchaudio_fbank_or_mfcc(w[None, :], 16000, n_bins=80)      # float32 [m,80]
        ref_cm = (ref_fb - mean) / std                                            # GlobalCMVN.__call__
        with data_utils.numpy_seed(1, 1, i):
            ref_sa = aug(ref_cm)
        # oracle restatement
        o_fb = O.kaldi_fbank(w)
        o_cm = O.global_cmvn(o_fb, mean, std)
        with O.numpy_seed(1, 1, i):
            o_sa, fm, tm = O.adaptive_specaugment(o_cm, return_masks=True)
        d_fb = np.abs(o_fb - ref_fb).max()
        d_sa = np.abs(o_sa - ref_sa).max()
        worst = max(worst, d_fb, d_sa)
        print("frontend utt %d: m=%d |fbank diff|=%.3g |specaug diff|=%.3g masks f=%s t=%d" % (
            i, ref_fb.shape[0], d_fb, d_sa, fm, len(tm)))
        assert d_fb < 2e-3 and d_sa < 2e-3, "oracle front end disagrees with the reference"
        out["wave_%d" % i] = w
        out["fbank_%d" % i] = ref_fb.astype(np.float32)
        out["final_%d" % i] = ref_sa.astype(np.float32)
        out["fmask_%d" % i] = np.array(fm, dtype=np.int32).reshape(-1, 2)
        out["tmask_%d" % i] = np.array(tm, dtype=np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(GOLDEN, "frontend.npz"), **out)
    print("frontend pinned (worst abs diff %.3g) -> tests/golden/frontend.npz" % worst)


def pin_ctc():
    import torch.nn.functional as F

    from oracle import ctc as O

    rs = np.random.RandomState(3)
    B, T, V, blank = 4, 23, 11, 0
    logits = torch.from_numpy(rs.randn(B, T, V).astype(np.float32) * 2.0).to(torch.bfloat16)
    in_lens = np.array([23, 17, 9, 3], dtype=np.int32)
    tgts = [rs.randint(1, V, size=u) for u in (7, 5, 4, 5)]  # last one infeasible (U > T)
    tgts[1][1] = tgts[1][0]  # repeated label
    u_max = max(len(t) for t in tgts)
    targets = np.zeros((B, u_max), dtype=np.int32)
    for b, t in enumerate(tgts):
        targets[b, :len(t)] = t
    tgt_lens = np.array([len(t) for t in tgts], dtype=np.int32)
    # the reference call (espresso/criterions/ctc_loss.py:85-94) on [T,B,V] fp32 log-probs
    x = logits.float().transpose(0, 1).contiguous().requires_grad_(True)
    lprobs = F.log_softmax(x, dim=-1)
    flat = torch.from_numpy(np.concatenate(tgts)).long()
    loss = F.ctc_loss(lprobs, flat, torch.from_numpy(in_lens).long(), torch.from_numpy(tgt_lens).long(),
                      blank=blank, reduction="none", zero_infinity=True)
    loss.sum().backward()
    ref_loss = loss.detach().numpy()
    ref_grad = x.grad.transpose(0, 1).numpy()  # [B,T,V]
    for b in range(B):
        nll, g = O.ctc_loss_and_grad(logits[b].float().numpy(), in_lens[b], tgts[b], blank)
        assert abs(nll - ref_loss[b]) < 1e-4 * max(1.0, abs(ref_loss[b])), (b, nll, ref_loss[b])
        assert np.abs(g - ref_grad[b]).max() < 1e-5, (b, np.abs(g - ref_grad[b]).max())
    print("ctc losses", ref_loss)
    np.savez_compressed(os.path.join(GOLDEN, "ctc.npz"), logits=logits.float().numpy(), in_lens=in_lens,
                        targets=targets, tgt_lens=tgt_lens, blank=blank, loss=ref_loss, grad=ref_grad)
    print("ctc pinned -> tests/golden/ctc.npz")


SECTIONS = {"frontend": pin_frontend, "ctc": pin_ctc}


def main(argv):
    refshim.activate()
    torch.manual_seed(0)
    os.makedirs(GOLDEN, exist_ok=True)
    for name in (argv or list(SECTIONS)):
        SECTIONS[name]()


if __name__ == "__main__":
    main(sys.argv[1:])
