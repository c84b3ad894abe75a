"""How well do a tensor-bound convolution and an HBM-bound BatchNorm pass share the GPU when issued on two streams?
Times R launches of each alone and then interleaved on two streams (distinct tensors, so nothing serialises them but the
hardware).  The answer bounds what pairing independent passes (trainer._fork_join) can win.  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

N, D, H, W, C = 8, 80, 96, 80, 64
dev = "cuda"
R = 10
x = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
y = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
g = torch.randn(N, D, H, W, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.05
wp, wpT = K.pack_conv3_weights(w)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
mean, invstd, scale, shift = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
x1 = torch.rand(N, D, H, W, device=dev)
w27 = torch.randn(C, 27, device=dev) * 0.1
b64 = torch.randn(C, device=dev)

s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def conv():
    K.conv3_igemm(x, wp)


def wgrad():
    K.conv3_wgrad(x, g)


def bn_fwd():
    K.bn_act_fwd(y, scale, shift, None, 0.2, 0)


def bn_bwd():
    K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0)


def stem():
    K.c1_to_cn(x1, w27, b64)


def run(fa, fb):
    """R launches of fa on s1 and (if given) R of fb on s2; returns elapsed ms from a common start to both done."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(R):
        if fa is not None:
            with torch.cuda.stream(s1):
                fa()
        if fb is not None:
            with torch.cuda.stream(s2):
                fb()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / R


for f in (conv, wgrad, bn_fwd, bn_bwd, stem):
    with torch.cuda.stream(s1):
        f()
    with torch.cuda.stream(s2):
        f()
torch.cuda.synchronize()
print(f"{R} launches each, ms per launch (pair): alone A, alone B, both on two streams, serial sum, overlap efficiency")
for na, fa, nb, fb in [("conv3 64->64", conv, "bn_act_fwd", bn_fwd), ("conv3 64->64", conv, "bn_act_bwd", bn_bwd),
                       ("wgrad 64x64", wgrad, "bn_act_fwd", bn_fwd), ("conv3 64->64", conv, "stem 1->64", stem),
                       ("conv3 64->64", conv, "conv3 64->64", conv), ("bn_act_fwd", bn_fwd, "bn_act_bwd", bn_bwd)]:
    run(fa, fb)
    ta, tb, tab = run(fa, None), run(None, fb), run(fa, fb)
    ideal = max(ta, tb)
    print(f"{na:14s} | {nb:12s}  A {ta:6.3f}  B {tb:6.3f}  both {tab:6.3f}  sum {ta + tb:6.3f}  "
          f"hidden {(ta + tb - tab) / min(ta, tb) * 100:5.1f} % of the shorter one")
