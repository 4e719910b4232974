"""GPU: speech_lstm (BASELINE configs[0] family) through cuDNN LSTMs + the native conv-front BatchNorm and the fused
label-smoothed cross-entropy, against the reference fixture."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_speech_lstm_vs_reference_fixture_on_gpu(golden_dir):
    from test_host_orchestration import _Task, _build_speech_lstm
    from espresso_b200 import lib
    from espresso_b200.criterions import LabelSmoothedCrossEntropyV2Criterion

    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "speech_lstm.npz"))
    m = _build_speech_lstm(g).finalize_(dev)
    crit = LabelSmoothedCrossEntropyV2Criterion(_Task(50), label_smoothing=0.1)
    sample = {"net_input": {"src_tokens": torch.from_numpy(g["feats"]).to(dev), "src_lengths": torch.from_numpy(g["lens"]).to(dev),
                            "prev_output_tokens": torch.from_numpy(g["prev_output_tokens"]).to(dev)},
              "target": torch.from_numpy(g["target"]).to(dev)}
    n0 = lib.launch_count()
    m.train()
    m.flat.zero_grad()
    loss, _, log = crit(m, sample)
    loss.backward()
    m.sync_torch_grads_()
    torch.cuda.synchronize()
    assert lib.launch_count() - n0 >= 10  # native BatchNorm / LS-CE kernels ran
    assert abs(loss.item() - float(g["loss"])) < 0.03 * float(g["loss"])
    worst = []
    for k in g.files:
        if k.startswith("grad.") and not ("pre_encoder.convolutions" in k and k.endswith(".bias")):
            refg = g[k]
            worst.append((np.linalg.norm(m.flat.grad(k[5:]).cpu().numpy() - refg) / max(np.linalg.norm(refg), 1e-3), k))
    worst.sort(reverse=True)
    assert worst[0][0] < 0.3, worst[:5]
    assert np.median([w[0] for w in worst]) < 0.06
    # the optimizer's flat buffer still backs the LSTM weights
    w = m.encoder.lstm[0].weight_ih_l0
    assert w.data_ptr() == m.flat.param("encoder.lstm.0.weight_ih_l0").data_ptr()
