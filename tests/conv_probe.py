"""Diagnostic probe for the tcgen05 convolution kernels (run on the GPU box).

Prints, per shape, the error of conv3_igemm / conv3_wgrad against torch fp32 on bf16-rounded inputs for
structured weight patterns (centre tap only, single off-centre tap, full) so a failure can be
attributed to the UMMA descriptors, the TMA halo coordinates or the epilogue.  Also times the kernels."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402
from sivae_b200 import kernels as K  # noqa: E402
from oracle import kernel_spec as S  # noqa: E402


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20)), float((a - b).abs().max())


def probe_fprop(n, d, h, w, ci, co, pattern):
    torch.manual_seed(0)
    x = torch.randn(n, d, h, w, ci, device="cuda").to(torch.bfloat16)
    wt = torch.randn(co, ci, 3, 3, 3, device="cuda") * (1.0 / (27 * ci) ** 0.5)
    if pattern == "centre":
        m = torch.zeros(27, device="cuda"); m[13] = 1
        wt = wt * m.view(1, 1, 3, 3, 3)
    elif pattern == "tap0":
        m = torch.zeros(27, device="cuda"); m[0] = 1
        wt = wt * m.view(1, 1, 3, 3, 3)
    elif pattern == "tap26":
        m = torch.zeros(27, device="cuda"); m[26] = 1
        wt = wt * m.view(1, 1, 3, 3, 3)
    wf, wd = K.pack_conv3_weights(wt)
    wf_s, wd_s = S.pack_conv3_weights(wt)
    pack_ok = torch.equal(wf, wf_s) and torch.equal(wd, wd_s)
    y = K.conv3_igemm(x, wf)
    torch.cuda.synchronize()
    ref = S.conv3_igemm(x, wf_s)
    e = rel_err(y, ref)
    return pack_ok, e


def probe_wgrad(n, d, h, w, ci, co):
    torch.manual_seed(1)
    x = torch.randn(n, d, h, w, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, d, h, w, co, device="cuda").to(torch.bfloat16)
    dw = K.conv3_wgrad(x, dy)
    torch.cuda.synchronize()
    ref = S.conv3_wgrad(x, dy)
    return rel_err(dw, ref)


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    K.device_check()
    print("device:", torch.cuda.get_device_name(0))
    shapes = [(1, 8, 8, 16, 64, 64), (2, 8, 12, 16, 64, 64), (1, 10, 12, 10, 128, 256), (1, 6, 8, 20, 256, 128),
              (1, 20, 24, 20, 64, 128)]
    for shp in shapes:
        for pat in ("centre", "tap0", "tap26", "full"):
            try:
                ok, e = probe_fprop(*shp, pat)
                print(f"fprop {shp} {pat:7s} pack_ok={ok} rel={e[0]:.3e} maxabs={e[1]:.3e}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"fprop {shp} {pat}: EXC {type(ex).__name__}: {ex}", flush=True)
                return 1
    for shp in shapes:
        try:
            e = probe_wgrad(*shp)
            print(f"wgrad {shp} rel={e[0]:.3e} maxabs={e[1]:.3e}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"wgrad {shp}: EXC {type(ex).__name__}: {ex}", flush=True)
            return 1
    # timing on the dominant layer shapes (B=8 would be 8x; use B=2 to stay quick)
    for (n, d, h, w, ci, co) in [(2, 80, 96, 80, 64, 64), (8, 40, 48, 40, 64, 128), (8, 20, 24, 20, 128, 256),
                                 (8, 10, 12, 10, 256, 256)]:
        x = torch.randn(n, d, h, w, ci, device="cuda").to(torch.bfloat16)
        dy = torch.randn(n, d, h, w, co, device="cuda").to(torch.bfloat16)
        wt = torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.02
        wf, wd = K.pack_conv3_weights(wt)
        fl = 2.0 * 27 * ci * co * n * d * h * w
        t = timeit(lambda: K.conv3_igemm(x, wf))
        print(f"time fprop {(n, d, h, w, ci, co)}: {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s", flush=True)
        t = timeit(lambda: K.conv3_wgrad(x, dy))
        print(f"time wgrad {(n, d, h, w, ci, co)}: {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
