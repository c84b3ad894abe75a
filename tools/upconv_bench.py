"""Times the upsample-folded convolution kernels in isolation (D4b shape by default)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

n, d, h, w, ci, co = [int(a) for a in (sys.argv[1:7] if len(sys.argv) >= 7 else (8, 40, 48, 40, 64, 64))]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 3
torch.manual_seed(0)
x = torch.randn(n, d, h, w, ci, device="cuda").to(torch.bfloat16)
dy = torch.randn(n, 2 * d, 2 * h, 2 * w, co, device="cuda").to(torch.bfloat16)
wt = torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.02
wup, wupT = K.pack_upconv3_weights(wt)
flops = 2.0 * 27 * ci * co * n * d * h * w * 8
for name, fn in (("fprop", lambda: K.upconv3_fprop(x, wup)), ("dgrad", lambda: K.upconv3_dgrad(dy, wupT)),
                 ("wgrad", lambda: K.upconv3_wgrad(x, dy))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name} {(n, d, h, w, ci, co)}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (ref-equiv)")
