"""Launches each hot kernel of the training step once (after one warm-up launch) at the headline full-resolution
size, for `ncu --set full` captures (profiles/*_ncu_*.md).  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

N, D, H, W, C = 8, 80, 96, 80, 64
dev = "cuda"
torch.manual_seed(0)
y = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
g = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
xlo = torch.randn(N, D // 2, H // 2, W // 2, C, device=dev).to(torch.bfloat16)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
wt = torch.randn(C, C, 3, 3, 3, device=dev) * 0.02
wf, wd = K.pack_conv3_weights(wt)
wup, wupT = K.pack_upconv3_weights(wt)
mean, invstd, scale, shift = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
x1 = torch.rand(N, D, H, W, device=dev)
w27 = torch.randn(C, 27, device=dev) * 0.1
b64 = torch.randn(C, device=dev)
b1 = torch.randn(1, device=dev)
targets = [
    lambda: K.conv3_igemm(y, wf),
    lambda: K.conv3_wgrad(y, g),
    lambda: K.upconv3_fprop(xlo, wup),
    lambda: K.upconv3_dgrad(g, wupT),
    lambda: K.upconv3_wgrad(xlo, g),
    lambda: K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5),
    lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 0),
    lambda: K.bn_act_fwd(y, scale, shift, None, 0.2, 1),
    lambda: K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0),
    lambda: K.c1_to_cn(x1, w27, b64),
    lambda: K.cn_to_c1(y, w27, b1, False, 1, None, 0.35, 1),
    lambda: K.wgrad_c1(y, x1, 27),
]
for rep in range(2):
    for fn in targets:
        fn()
    torch.cuda.synchronize()
print("done")
