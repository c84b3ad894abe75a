"""ncu target: one launch of each Linear-head kernel at the BASELINE config-2 shapes (fc 38400x1200, dfc 600x38400, B=4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sivae_b200 import kernels as K  # noqa: E402

dev = "cuda"
for k, j in ((38400, 1200), (600, 38400)):
    b = 4
    x = torch.randn(b, k, device=dev)
    w = torch.randn(j, k, device=dev) / k ** 0.5
    bias = torch.randn(j, device=dev)
    dy = torch.randn(b, j, device=dev)
    for _ in range(2):
        K.linear_fwd(x, w, bias)
        K.linear_dgrad(dy, w)
        K.linear_wgrad(x, dy)
    torch.cuda.synchronize()
print("ok")
