"""One launch of every memory-bound kernel at the headline full-resolution size (8 x 80x96x80 x 64 ch), for
`ncu --set full` (tools/gpu/ncu_pointwise.sh): DRAM bytes / throughput per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

N, D, H, W, C = 8, 80, 96, 80, 64
dev = "cuda"
y = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
g = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
gp = torch.randn(N, D // 2, H // 2, W // 2, C, device=dev).to(torch.bfloat16)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
mean, invstd, scale, shift = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
x1 = torch.rand(N, D, H, W, device=dev)
bits = torch.empty(y.numel() // 8, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
K.bn_act_fwd(y, scale, shift, None, 0.2, 0)                                        # bn_act_plain_fwd_kernel
K.bn_act_fwd(y, scale, shift, g, 0.2, 0)                                           # bn_act_fwd_kernel<0>, residual
K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.35, 1, keep_bits=bits)         # + Philox dropout, keep-bit store
K.bn_act_fwd(y, scale, shift, None, 0.2, 1)                                        # bn_act_fwd_kernel<1>, AvgPool3d(2)
K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0)                        # bn_bwd_{reduce,apply}_plain_kernel<0>
K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0, None, 0.35, 1, keep_bits=bits)   # ..._plain_kernel<2>
K.bn_act_bwd(gp, y, None, mean, invstd, gamma, beta, 0.2, 1)                       # bn_act_bwd_{reduce,apply}_kernel<1>
K.mse_persample_fwd(x1.view(N, -1), x1.view(N, -1))
torch.cuda.synchronize()
print("done")
