"""Per-tensor gradient comparison at the full-size parity configuration: CUDA path vs the host path over oracle/ops_ref
(both bf16).  Debug aid for tests/test_gpu_encoder.py::test_full_size_encoder_value_parity.
    python profiles/debug_fullsize.py gpu|cpu|cmp"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
mode = sys.argv[1]
if mode in ("gpu", "cpu"):
    if mode == "cpu":
        from espresso_b200 import ops
        from oracle import ops_ref

        for name in dir(ops_ref):
            if not name.startswith("_") and callable(getattr(ops_ref, name)) and hasattr(ops, name):
                setattr(ops, name, getattr(ops_ref, name))
    import fullsize_util as F
    from oracle.fullsize import fullsize_cotangent, fullsize_inputs

    dev = torch.device("cuda:0" if mode == "gpu" else "cpu")
    g = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_conformer.npz"))
    feats, lens, tgt = fullsize_inputs()
    m = F.build_model(dev)
    m.train()
    m.flat.zero_grad()
    net = m(src_tokens=torch.from_numpy(feats).to(dev), src_lengths=torch.from_numpy(lens).to(dev), src_lengths_cpu=torch.from_numpy(lens))
    out_full = net["b200_out"]
    G = torch.from_numpy(fullsize_cotangent(g["out_lens"].tolist()))
    Gp = torch.zeros(out_full.shape, dtype=out_full.dtype, device=dev)
    Gp[..., : G.shape[-1]] = G.to(dev)
    out_full.backward(Gp)
    m.encoder.sync_torch_grads_()
    torch.save({n: m.flat.grad(n).float().cpu() for n in m.flat.names}, "/tmp/grads_%s.pt" % mode)
    gn = float(torch.sqrt((m.flat.grads.double() ** 2).sum()))
    print(mode, "|grad| =", gn, " fixture fp32", float(g["gnorm_fp32"]), "bf16", float(g["gnorm_bf16"]))
else:
    a, b = torch.load("/tmp/grads_gpu.pt"), torch.load("/tmp/grads_cpu.pt")
    rows = []
    for n in a:
        d = (a[n] - b[n]).norm().item()
        rows.append((d / max(b[n].norm().item(), 1e-12), d, b[n].norm().item(), a[n].norm().item(), n))
    rows.sort(reverse=True)
    for r in rows[:25]:
        print("rel %.3e  |diff| %.4e  |cpu| %.4e  |gpu| %.4e  %s" % r)
    rows.sort(key=lambda r: -r[1])
    print("largest absolute differences:")
    for r in rows[:10]:
        print("rel %.3e  |diff| %.4e  |cpu| %.4e  |gpu| %.4e  %s" % r)
