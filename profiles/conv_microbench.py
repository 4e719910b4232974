"""Device time of the native conv-front kernels at the bench workload's shapes (26 000 frames x 80 bins per batch) next to the
library convolution (cuDNN through torch, channels-last bf16) on the same tensors.

    python profiles/conv_microbench.py            # on a B200
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espresso_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, FB = 24, 1083, 80
LAYERS = [(1, 64, (1, 1)), (64, 64, (2, 2)), (64, 128, (1, 1)), (128, 128, (2, 2))]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


t, f = T, FB
for cin, cout, st in LAYERS:
    to, fo = (t + st[0] - 1) // st[0], (f + st[1] - 1) // st[1]
    x = torch.randn(B, t, f, cin, device=dev).bfloat16()
    if cin == 1:
        x = x.reshape(B, t, f)
    w = (torch.randn(cout, 3, 3, cin, device=dev) * (9 * cin) ** -0.5).bfloat16()
    dy = torch.randn(B, to, fo, cout, device=dev).bfloat16()
    dw = torch.zeros(cout, 3, 3, cin, device=dev)
    flops = 2.0 * B * to * fo * cout * 9 * cin
    xl = (x.reshape(B, t, f, cin)).permute(0, 3, 1, 2)  # NCHW view of channels-last memory
    wl = w.permute(0, 3, 1, 2)
    dyl = dy.permute(0, 3, 1, 2)
    rows = [("fwd", lambda: ops.conv3x3_fwd(x, w, st), lambda: F.conv2d(xl, wl, None, st, (1, 1))),
            ("wgrad", lambda: ops.conv3x3_wgrad(dy, x, dw, st),
             lambda: torch.ops.aten.convolution_backward(dyl, xl, wl, None, st, (1, 1), (1, 1), False, (0, 0), 1, (False, True, False)))]
    if cin > 1:
        rows.append(("dgrad", lambda: ops.conv3x3_dgrad(dy, w, (B, t, f, cin), st),
                     lambda: torch.ops.aten.convolution_backward(dyl, xl, wl, None, st, (1, 1), (1, 1), False, (0, 0), 1, (True, False, False))))
    for name, mine, lib in rows:
        a, b = timeit(mine), timeit(lib)
        print("conv %3d->%3d stride %s  %-5s  native %7.1f us (%6.1f TFLOP/s)   library %7.1f us   [%d x %d x %d -> %d x %d]"
              % (cin, cout, st, name, a, flops / a / 1e6, b, B, t, f, to, fo))
    t, f = to, fo
