"""GPU parity of the native 3x3 convolutions of the conv front end (espresso/modules/speech_convolutions.py:78-102) against
the fp32 torch restatement in oracle/ops_ref.py: forward, input gradient (stride-1 and the four parity classes of stride 2)
and weight gradient, incl. odd sizes (partial boxes, out-of-bounds taps) and the one-input-channel first layer."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # B, T, F, Cin, Cout, (st, sf)
    (2, 37, 40, 64, 64, (2, 2)),
    (3, 50, 20, 64, 128, (1, 1)),
    (2, 33, 21, 128, 128, (2, 2)),
    (1, 130, 40, 64, 64, (1, 1)),
    (2, 40, 24, 128, 64, (2, 1)),
    (2, 19, 83, 64, 64, (1, 2)),
]


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_matches_fp32_reference(case):
    from espresso_b200 import ops
    from oracle import ops_ref

    B, T, F_, Cin, Cout, stride = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(sum(case[:5]))
    x = torch.randn(B, T, F_, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) * (9 * Cin) ** -0.5).to(dev).bfloat16()
    y = ops.conv3x3_fwd(x, w, stride)
    y_ref = ops_ref.conv3x3_fwd(x, w, stride)
    assert y.shape == y_ref.shape
    assert _rel(y, y_ref) < 1e-2, "forward"
    dy = torch.randn(y.shape, generator=g).to(dev).bfloat16()
    dx = ops.conv3x3_dgrad(dy, w, x.shape, stride)
    dx_ref = ops_ref.conv3x3_dgrad(dy, w, x.shape, stride)
    assert _rel(dx, dx_ref) < 1e-2, "input gradient"
    dw = torch.zeros(Cout, 3, 3, Cin, device=dev)
    dw_ref = torch.zeros(Cout, 3, 3, Cin, device=dev)
    ops.conv3x3_wgrad(dy, x, dw, stride)
    ops.conv3x3_wgrad(dy, x, dw, stride)  # accumulates
    ops_ref.conv3x3_wgrad(dy, x, dw_ref, stride)
    assert _rel(dw, 2 * dw_ref) < 2e-3, "weight gradient"


@pytest.mark.parametrize("case", [(2, 45, 80, 64, (1, 1)), (3, 31, 83, 32, (2, 2)), (1, 400, 80, 64, (1, 1))])
def test_conv3x3_single_input_channel(case):
    from espresso_b200 import ops
    from oracle import ops_ref

    B, T, F_, Cout, stride = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(T)
    x = torch.randn(B, T, F_, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, 3, 3, generator=g) / 3).to(dev).bfloat16()
    y = ops.conv3x3_fwd(x, w, stride)
    y_ref = ops_ref.conv3x3_fwd(x, w, stride)
    assert y.shape == y_ref.shape and _rel(y, y_ref) < 1e-2
    dy = torch.randn(y.shape, generator=g).to(dev).bfloat16()
    dw = torch.zeros(Cout, 3, 3, device=dev)
    dw_ref = torch.zeros(Cout, 3, 3, device=dev)
    ops.conv3x3_wgrad(dy, x, dw, stride)
    ops_ref.conv3x3_wgrad(dy, x, dw_ref, stride)
    assert _rel(dw, dw_ref) < 2e-3
