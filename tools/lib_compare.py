"""Two builds of libsivae.so on the thin convolution kernels, same inputs: bitwise comparison of the outputs.

    python tools/lib_compare.py OLD.so [NEW.so]       (NEW defaults to the in-tree library; run on the GPU box)

Used in round 2 to separate a kernel restructuring that must not change results (conv3_to1: converter / gather warp
groups) from one that may move single roundings (c1_to_c64: bias through a K slot of the GEMM)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

old = K.load_library(os.path.abspath(sys.argv[1]))
new = K.load_library(os.path.abspath(sys.argv[2])) if len(sys.argv) > 2 else K.load_library()
torch.manual_seed(0)
N, D, H, W, C = 8, 80, 96, 80, 64
x = torch.randn(N, D, H, W, C, device="cuda").to(torch.bfloat16)
x1 = torch.rand(N, D, H, W, device="cuda")
w27 = torch.randn(C, 27, device="cuda") * 0.1
b64 = torch.randn(C, device="cuda")
b1 = torch.randn(1, device="cuda")
x1_sparse = x1 * (torch.rand_like(x1) > 0.6)


def both(fn):
    K._lib = old
    a = fn()
    torch.cuda.synchronize()
    K._lib = new
    b = fn()
    torch.cuda.synchronize()
    return a, b


cases = [("cn_to_c1 relu + philox dropout", lambda: K.cn_to_c1(x, w27, b1, False, 1, None, 0.35, 1234)),
         ("cn_to_c1 plain, flipped taps", lambda: K.cn_to_c1(x, w27, None, True, 0)),
         ("cn_to_c1 2x37x45x51", lambda: K.cn_to_c1(x[:2, :37, :45, :51].contiguous(), w27, b1, False, 1, None, 0.35, 99)),
         ("c1_to_cn with bias", lambda: K.c1_to_cn(x1, w27, b64)),
         ("c1_to_cn no bias, flipped taps", lambda: K.c1_to_cn(x1, w27, None, True)),
         ("c1_to_cn no bias, 60 % zero input, negative filters",
          lambda: K.c1_to_cn(x1_sparse, -w27.abs(), None, True)),
         ("wgrad_c1", lambda: K.wgrad_c1(x, x1, 27)[0])]
for name, fn in cases:
    a, b = both(fn)
    for _ in range(3):                      # the new build must also agree with itself run to run
        c = fn()
        torch.cuda.synchronize()
        assert torch.equal(b, c), f"{name}: new build not deterministic"
    af, bf = a.float(), b.float()
    ne = a.view(torch.int16) != b.view(torch.int16) if a.dtype == torch.bfloat16 else af != bf   # -0 vs +0 counts
    worst = float(((af - bf).abs() / af.abs().clamp_min(1e-6))[ne].max()) if bool(ne.any()) else 0.0
    print(f"{name:34s} bitwise equal: {bool(torch.equal(a, b))!s:5s}  elements differing {100 * ne.float().mean().item():.4f} %  "
          f"worst relative difference {worst:.2e}")
