"""Achieved HBM bandwidth of the memory-bound kernels at the headline full-resolution size (8 x 80x96x80 x 64 ch).
Algorithmic bytes = tensors that must be read/written once.  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

N, D, H, W, C = 8, 80, 96, 80, 64
dev = "cuda"
T = N * D * H * W * C * 2 / 1e9   # GB of one bf16 full-res tensor
F1 = N * D * H * W * 4 / 1e9      # GB of one fp32 1-channel tensor


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = 0.0
    for _ in range(iters):
        big.zero_()                       # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters


y = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
g = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
gp = torch.randn(N, D // 2, H // 2, W // 2, C, device=dev).to(torch.bfloat16)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
mean, invstd, scale, shift = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
x1 = torch.rand(N, D, H, W, device=dev)
w27 = torch.randn(C, 27, device=dev) * 0.1
b64 = torch.randn(C, device=dev)
b1 = torch.randn(1, device=dev)
bits = torch.empty(y.numel() // 8, dtype=torch.uint8, device=dev)
K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.35, 1, keep_bits=bits)
rows = [
    ("bn_train_coeffs (stats)", T, lambda: K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)),
    ("bn_act_fwd none", 2 * T, lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 0)),
    ("bn_act_fwd none + philox dropout", 2 * T, lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.35, 1)),
    ("bn_act_fwd none + philox + keep-bit store", 2 * T + T / 16,
     lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.35, 1, keep_bits=bits)),
    ("bn_act_fwd + residual (2nd BN of a block)", 3 * T, lambda: K.bn_act_fwd(y, scale, shift, g, 0.2, 0)),
    ("bn_act_fwd avgpool", T + T / 8, lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 1)),
    ("bn_act_bwd + residual gradient", 6 * T,
     lambda: K.bn_act_bwd(g, y, g, mean, invstd, gamma, beta, 0.2, 0, need_dres=True)),
    ("bn_act_bwd none (reduce+apply)", 5 * T, lambda: K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0)),
    ("bn_act_bwd none + philox", 5 * T, lambda: K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0, None, 0.35, 1)),
    ("bn_act_bwd none + keep bits (as the step runs it)", 5 * T + 2 * T / 16,
     lambda: K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0, None, 0.35, 1, keep_bits=bits)),
    ("bn_act_bwd avgpool (reduce+apply)", 3 * T + 2 * T / 8, lambda: K.bn_act_bwd(gp, y, None, mean, invstd, gamma, beta, 0.2, 1)),
    ("c1_to_cn 27 taps (stem fwd / tail dgrad)", T + F1, lambda: K.c1_to_cn(x1, w27, b64)),
    ("cn_to_c1 27 taps tcgen05 (tail fwd)", T + F1, lambda: K.cn_to_c1(y, w27, b1, False, 1, None, 0.35, 1)),
    ("wgrad_c1 27 taps", T + F1, lambda: K.wgrad_c1(y, x1, 27)),
    ("mse_persample_fwd", 2 * F1, lambda: K.mse_persample_fwd(x1.view(N, -1), x1.view(N, -1))),
    ("reparam_draw_fwd (in-kernel sampler, 8x1200)", 4 * 8 * 1200 * 4 / 1e9,
     lambda: K.reparam_draw_fwd(x1.view(-1)[:9600].contiguous(), x1.view(-1)[:9600].contiguous(), 7)),
]
print(f"one full-res bf16 tensor = {T:.3f} GB")
for name, gb, fn in rows:
    ms = timeit(fn)
    print(f"{name:44s} {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s  ({gb:.2f} GB algorithmic)")
