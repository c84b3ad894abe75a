"""HBM roofline of the Linear-head kernels (csrc/linear.cu) at the BASELINE config-2 shapes:
fc = Linear(38400, 1200), dfc = Linear(600, 38400), batch 4 and 8.  CUDA events around a graph of 20 launches each; the 184 MB
weight matrix is larger than the 126 MB L2, so every launch streams it from HBM.  Run on the GPU box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sivae_b200 import kernels as K  # noqa: E402

peak = 6546.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass
dev = "cuda"


def timed(fn, n=20):
    """n launches captured in one CUDA graph (as the training step runs them), so the host-side ctypes / allocator
    cost per call (~20 us, comparable to these kernels) does not gate the device."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"HBM copy peak (MEASURED_PEAKS.json): {peak:.0f} GB/s")
print("| kernel | B | K | J | ms | algorithmic bytes | GB/s | frac of peak |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for name, k, j in (("fc", 38400, 1200), ("dfc", 600, 38400)):
    for b in (4, 8):
        x = torch.randn(b, k, device=dev)
        w = torch.randn(j, k, device=dev) / k ** 0.5
        bias = torch.randn(j, device=dev)
        dy = torch.randn(b, j, device=dev)
        wb = w.numel() * 4
        io = (x.numel() + dy.numel()) * 4
        for kern, fn, nbytes in ((f"{name} linear_fwd", lambda: K.linear_fwd(x, w, bias), wb + io),
                                 (f"{name} linear_dgrad", lambda: K.linear_dgrad(dy, w), wb + io),
                                 (f"{name} linear_wgrad", lambda: K.linear_wgrad(x, dy), wb + io)):
            ms = timed(fn)
            gbs = nbytes / ms / 1e6
            print(f"| {kern} | {b} | {k} | {j} | {ms:.4f} | {nbytes / 1e6:.1f} MB | {gbs:.0f} | {gbs / peak:.2f} |")
