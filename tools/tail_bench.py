"""CUDA-event timing of the decoder-end backward at the headline size: the unfused three launches (tail input gradient ->
BatchNorm-backward reduce -> apply) vs sivae_tail_dgrad_bn_bwd (reduce and apply passes that recompute the gradient)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

N, D, H, W = 8, 80, 96, 80
dev = "cuda"
torch.manual_seed(0)
y = torch.randn(N, D, H, W, 64, device=dev).to(torch.bfloat16)
dy1 = torch.randn(N, D, H, W, device=dev)
wt = torch.randn(64, 27, device=dev) * 0.1
gamma, beta = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
mean, invstd, _, _ = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def unfused():
    g = K.c1_to_cn(dy1, wt, None, flip=True)
    return K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0)


t_conv = timeit(lambda: K.c1_to_cn(dy1, wt, None, flip=True))
g = K.c1_to_cn(dy1, wt, None, flip=True)
t_bn = timeit(lambda: K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0))
t_unf = timeit(unfused)
t_fus = timeit(lambda: K.tail_dgrad_bn_bwd(dy1, wt, y, mean, invstd, gamma, beta, 0.2))
print(f"tail input gradient (c1_to_c64<0>): {t_conv:.3f} ms   BatchNorm backward (reduce+finalize+apply): {t_bn:.3f} ms   "
      f"unfused total: {t_unf:.3f} ms   fused (c1_to_c64<1> + finalize + c1_to_c64<2>): {t_fus:.3f} ms")
