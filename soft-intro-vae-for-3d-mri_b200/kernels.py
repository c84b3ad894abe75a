"""Tensor-level bindings of libsivae.so (the C ABI declared in include/sivae.h).

Every function here is one call into the shared library (one or a few kernel launches on the
current torch CUDA stream).  There is NO fallback: if the library is missing, fails to load,
or a tensor is not on a CUDA device, the call raises.  torch is used only for device memory
(``torch.empty``) and the stream handle.

Layout conventions: activations are NDHWC bf16 tensors of shape [N, D, H, W, C]; one-channel model
inputs / outputs / latents are fp32 [N, D, H, W].
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIVAE_LIB") or os.path.join(_HERE, "libsivae.so")   # SIVAE_LIB: A/B another build

RESAMPLE_NONE, RESAMPLE_AVGPOOL2, RESAMPLE_UPSAMPLE2 = 0, 1, 2

_lib = None
_vp, _i, _ll, _f, _sz, _u64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                               ctypes.c_size_t, ctypes.c_ulonglong)

# name -> (restype, argtypes); mirrors include/sivae.h one to one
_SIGNATURES = {
    "sivae_last_error": (ctypes.c_char_p, []),
    "sivae_abi_version": (_i, []),
    "sivae_device_check": (_i, []),
    "sivae_launch_count": (_ll, []),
    "sivae_set_seed_counter": (_i, [_vp]),
    "sivae_advance_seed_counter": (_i, [_vp]),
    "sivae_pack_conv3_weights": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "sivae_conv3_igemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_conv3_igemm_bn": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp,
                                  _vp, _vp, _sz, _vp]),
    "sivae_conv3_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "sivae_conv3_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_pack_upconv3_weights": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "sivae_upconv3_fprop": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_upconv3_fprop_bn": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp,
                                    _vp, _vp, _sz, _vp]),
    "sivae_upconv3_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_upconv3_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "sivae_upconv3_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_conv3_to1_workspace_bytes": (_sz, [_i]),
    "sivae_conv3_to1": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _u64, _vp, _sz, _vp]),
    "sivae_bn_workspace_bytes": (_sz, [_i]),
    "sivae_bn_train_coeffs": (_i, [_vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sivae_bn_train_act_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp,
                                    _vp, _vp, _sz, _vp]),
    "sivae_bn_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _f, _u64, _vp]),
    "sivae_bn_act_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i,
                              _vp, _f, _u64, _vp, _sz, _vp]),
    "sivae_c1_to_cn": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "sivae_c1_to_c64_workspace_bytes": (_sz, []),
    "sivae_c1_to_c64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sivae_c1_to_c64_bn": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp,
                                _vp, _sz, _vp, _sz, _vp]),
    "sivae_cn_to_c1": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _u64, _vp]),
    "sivae_wgrad_c1_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "sivae_wgrad_c1": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sivae_wgrad_c64_workspace_bytes": (_sz, []),
    "sivae_wgrad_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sivae_relu_drop_bwd": (_i, [_vp, _vp, _vp, _ll, _f, _vp]),
    "sivae_reparam_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _ll, _vp]),
    "sivae_reparam_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _ll, _i, _vp]),
    "sivae_reparam_draw_fwd": (_i, [_vp, _vp, _vp, _vp, _ll, _u64, _vp]),
    "sivae_kl_persample_fwd": (_i, [_vp, _vp, _vp, _i, _ll, _vp]),
    "sivae_kl_persample_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp]),
    "sivae_mse_workspace_bytes": (_sz, [_i, _ll]),
    "sivae_mse_persample_fwd": (_i, [_vp, _vp, _vp, _i, _ll, _vp, _sz, _vp]),
    "sivae_mse_persample_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _vp]),
    "sivae_adam_step": (_i, [_vp, _i, _vp, _f, _f, _f, _vp, _vp]),
    "sivae_intro_loss_e_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _vp, _vp]),
    "sivae_intro_loss_e_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sivae_intro_loss_d_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _vp, _vp]),
    "sivae_intro_loss_d_bwd": (_i, [_vp, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sivae_volume_stats_workspace_bytes": (_sz, [_i]),
    "sivae_volume_stats": (_i, [_vp, _i, _ll, _vp, _vp, _sz, _vp]),
    "sivae_preprocess_clip_minmax": (_i, [_vp, _vp, _i, _ll, _f, _vp, _vp, _sz, _vp]),
    "sivae_affine_resample": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sivae_linear_workspace_bytes": (_sz, [_i, _i, _i]),
    "sivae_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sivae_linear_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "sivae_linear_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "sivae_ndhwc_to_flat": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sivae_flat_to_ndhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sivae_add_act_fwd": (_i, [_vp, _vp, _vp, _ll, _f, _vp]),
    "sivae_add_act_bwd": (_i, [_vp, _vp, _vp, _ll, _f, _vp]),
    "sivae_similarity_workspace_bytes": (_sz, [_i, _i]),
    "sivae_similarity_topk": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "sivae_ncdhw_f32_to_ndhwc_bf16": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
    "sivae_ndhwc_bf16_to_ncdhw_f32": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class SivaeError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen libsivae.so and declare the prototypes.  Raises SivaeError when it is absent
    (build it with ``python __graft_entry__.py`` or ``make -C .../csrc``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise SivaeError(f"{p} not found: the CUDA extension is not built (run `python __graft_entry__.py`); "
                         "there is no CPU or PyTorch fallback for the hot path")
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sivae_abi_version() != 1:
        raise SivaeError("libsivae.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def _L():
    return _lib if _lib is not None else load_library()


def _check(rc: int, what: str):
    if rc != 0:
        raise SivaeError(f"{what} failed ({rc}): {_L().sivae_last_error().decode()}")


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise SivaeError(f"{name}: expected a CUDA tensor (no CPU fallback exists for the hot path)")
    if t.dtype != dtype:
        raise SivaeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise SivaeError(f"{name}: expected a contiguous tensor")
    return t


def launch_count() -> int:
    return int(_L().sivae_launch_count())


def device_check():
    _check(_L().sivae_device_check(), "sivae_device_check")


class KernelTimer:
    """Optional CUDA-event timing of individual C-ABI calls on their launch stream (used by bench.py to
    report the dominant kernel's achieved FLOP/s inside the timed region).  Inactive unless entered."""

    active = None

    def __init__(self, names=("conv3_igemm", "conv3_wgrad", "conv3_to1", "upconv3_fprop", "upconv3_dgrad", "upconv3_wgrad")):
        self.names = set(names)
        self.records = []   # (name, work, start_event, end_event)

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *a):
        KernelTimer.active = None

    def summary(self):
        """-> {name: dict(launches, ms, work)}; call after torch.cuda.synchronize()."""
        out = {}
        for name, work, e0, e1 in self.records:
            d = out.setdefault(name, dict(launches=0, ms=0.0, work=0.0, by_shape={}))
            ms = e0.elapsed_time(e1)
            d["launches"] += 1
            d["ms"] += ms
            d["work"] += work[0]
            s = d["by_shape"].setdefault(work[1], dict(launches=0, ms=0.0, work=0.0))
            s["launches"] += 1
            s["ms"] += ms
            s["work"] += work[0]
        return out


def _timed(name, work, fn):
    t = KernelTimer.active
    if t is None or name not in t.names or torch.cuda.is_current_stream_capturing():
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    t.records.append((name, work, e0, e1))
    return r


def set_seed_counter(counter: Optional[torch.Tensor]):
    """Register (or with None: unregister) the device-resident int64/uint64 dropout epoch counter."""
    if counter is not None:
        _req(counter, torch.int64, "counter")
    _check(_L().sivae_set_seed_counter(_p(counter)), "sivae_set_seed_counter")


def advance_seed_counter(device):
    _check(_L().sivae_advance_seed_counter(ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
           "sivae_advance_seed_counter")


_ws_cache = {}


def _workspace(device, nbytes: int, tag: str) -> torch.Tensor:
    """Per-(device, stream, tag) scratch buffer, grown on demand (caller-owned workspace of the C ABI)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ----------------------------------------------------------------------------------------------
# 3x3x3 convolutions on tcgen05
# ----------------------------------------------------------------------------------------------
def pack_conv3_weights(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [Co,Ci,3,3,3] -> (wf bf16 [27,Co,Ci], wd bf16 [27,Ci,Co] flipped+transposed)."""
    _req(w, torch.float32, "weight")
    co, ci = w.shape[0], w.shape[1]
    wf = torch.empty(27, co, ci, dtype=torch.bfloat16, device=w.device)
    wd = torch.empty(27, ci, co, dtype=torch.bfloat16, device=w.device)
    _check(_L().sivae_pack_conv3_weights(_p(w), co, ci, _p(wf), _p(wd), _stream(w)), "sivae_pack_conv3_weights")
    return wf, wd


def conv3_igemm(x: torch.Tensor, wpack: torch.Tensor) -> torch.Tensor:
    """y[n,d,h,w,co] = sum_{tap,ci} x[n,d+kd-1,h+kh-1,w+kw-1,ci] * wpack[tap,co,ci]."""
    _req(x, torch.bfloat16, "x")
    _req(wpack, torch.bfloat16, "wpack")
    n, d, h, w, ci = x.shape
    co = wpack.shape[1]
    assert wpack.shape == (27, co, ci), (wpack.shape, ci)
    y = torch.empty(n, d, h, w, co, dtype=torch.bfloat16, device=x.device)
    flops = 2.0 * 27 * ci * co * n * d * h * w
    lib = _L()
    _timed("conv3_igemm", (flops, (n, d, h, w, ci, co)),
           lambda: _check(lib.sivae_conv3_igemm(_p(x), _p(wpack), _p(y), n, d, h, w, ci, co, _stream(x)),
                          "sivae_conv3_igemm"))
    return y


def conv3_igemm_bn(x: torch.Tensor, wpack: torch.Tensor, gamma, beta, running_mean, running_var, num_batches_tracked,
                   momentum: float, eps: float):
    """y = conv3(x), plus the train-mode BatchNorm coefficients of y: -> (y, mean, invstd, scale, shift).  The channel
    sums come from the convolution's epilogue when the persistent kernel runs (no extra pass over y)."""
    _req(x, torch.bfloat16, "x")
    _req(wpack, torch.bfloat16, "wpack")
    n, d, h, w, ci = x.shape
    taps, co, ci2 = wpack.shape
    assert taps == 27 and ci2 == ci
    lib = _L()
    y = torch.empty(n, d, h, w, co, dtype=torch.bfloat16, device=x.device)
    ws = _workspace(x.device, lib.sivae_bn_workspace_bytes(co), "bn")
    coef = torch.empty(4, co, dtype=torch.float32, device=x.device)
    if num_batches_tracked is not None:
        _req(num_batches_tracked, torch.int64, "num_batches_tracked")
    flops = 2.0 * 27 * ci * co * n * d * h * w
    _timed("conv3_igemm", (flops, (n, d, h, w, ci, co)),
           lambda: _check(lib.sivae_conv3_igemm_bn(_p(x), _p(wpack), _p(y), n, d, h, w, ci, co, _p(gamma), _p(beta),
                                                   _p(running_mean), _p(running_var), _p(num_batches_tracked), momentum,
                                                   eps, _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]), _p(ws),
                                                   ws.numel(), _stream(x)), "sivae_conv3_igemm_bn"))
    return y, coef[0], coef[1], coef[2], coef[3]


def conv3_wgrad(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """dw[co,ci,kd,kh,kw] = sum_v dy[v,co] * x[v+tap,ci]   (fp32, torch weight layout)."""
    _req(x, torch.bfloat16, "x")
    _req(dy, torch.bfloat16, "dy")
    n, d, h, w, ci = x.shape
    co = dy.shape[-1]
    assert dy.shape[:4] == x.shape[:4]
    lib = _L()
    nbytes = lib.sivae_conv3_wgrad_workspace_bytes(n, d, h, w, ci, co)
    ws = _workspace(x.device, nbytes, "wgrad")
    dw = torch.empty(co, ci, 3, 3, 3, dtype=torch.float32, device=x.device)
    flops = 2.0 * 27 * ci * co * n * d * h * w
    _timed("conv3_wgrad", (flops, (n, d, h, w, ci, co)),
           lambda: _check(lib.sivae_conv3_wgrad(_p(x), _p(dy), _p(dw), _p(ws), ws.numel(), n, d, h, w, ci, co,
                                                _stream(x)), "sivae_conv3_wgrad"))
    return dw


# ---- nearest-Upsample(2) folded into the 3x3x3 convolution (8 output parities x 8 pre-summed taps) ----
def pack_upconv3_weights(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [Co,Ci,3,3,3] -> (wup bf16 [64,Co,Ci], wupT bf16 [64,Ci,Co]); index = parity*8 + abc."""
    _req(w, torch.float32, "weight")
    co, ci = w.shape[0], w.shape[1]
    wup = torch.empty(64, co, ci, dtype=torch.bfloat16, device=w.device)
    wupT = torch.empty(64, ci, co, dtype=torch.bfloat16, device=w.device)
    _check(_L().sivae_pack_upconv3_weights(_p(w), co, ci, _p(wup), _p(wupT), _stream(w)), "sivae_pack_upconv3_weights")
    return wup, wupT


def upconv3_fprop(x_lo: torch.Tensor, wup: torch.Tensor) -> torch.Tensor:
    """y_hi = conv3(upsample2(x_lo), w) without materialising the upsampled tensor."""
    _req(x_lo, torch.bfloat16, "x_lo")
    _req(wup, torch.bfloat16, "wup")
    n, d, h, w, ci = x_lo.shape
    co = wup.shape[1]
    assert wup.shape == (64, co, ci)
    y = torch.empty(n, 2 * d, 2 * h, 2 * w, co, dtype=torch.bfloat16, device=x_lo.device)
    flops = 2.0 * 27 * ci * co * n * d * h * w * 8      # reference-equivalent (the kernel executes 8/27 of it)
    _timed("upconv3_fprop", (flops, (n, d, h, w, ci, co)),
           lambda: _check(_L().sivae_upconv3_fprop(_p(x_lo), _p(wup), _p(y), n, d, h, w, ci, co, _stream(x_lo)),
                          "sivae_upconv3_fprop"))
    return y


def upconv3_fprop_bn(x_lo: torch.Tensor, wup: torch.Tensor, gamma, beta, running_mean, running_var,
                     num_batches_tracked, momentum: float, eps: float):
    """upconv3_fprop + train-mode BatchNorm coefficients of the output: -> (y_hi, mean, invstd, scale, shift)."""
    _req(x_lo, torch.bfloat16, "x_lo")
    _req(wup, torch.bfloat16, "wup")
    n, d, h, w, ci = x_lo.shape
    co = wup.shape[1]
    assert wup.shape == (64, co, ci)
    lib = _L()
    y = torch.empty(n, 2 * d, 2 * h, 2 * w, co, dtype=torch.bfloat16, device=x_lo.device)
    ws = _workspace(x_lo.device, lib.sivae_bn_workspace_bytes(co), "bn")
    coef = torch.empty(4, co, dtype=torch.float32, device=x_lo.device)
    if num_batches_tracked is not None:
        _req(num_batches_tracked, torch.int64, "num_batches_tracked")
    flops = 2.0 * 27 * ci * co * n * d * h * w * 8
    _timed("upconv3_fprop", (flops, (n, d, h, w, ci, co)),
           lambda: _check(lib.sivae_upconv3_fprop_bn(_p(x_lo), _p(wup), _p(y), n, d, h, w, ci, co, _p(gamma), _p(beta),
                                                     _p(running_mean), _p(running_var), _p(num_batches_tracked),
                                                     momentum, eps, _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]),
                                                     _p(ws), ws.numel(), _stream(x_lo)), "sivae_upconv3_fprop_bn"))
    return y, coef[0], coef[1], coef[2], coef[3]


def upconv3_dgrad(dy_hi: torch.Tensor, wupT: torch.Tensor) -> torch.Tensor:
    """dx_lo = upsample2^T(conv3^T(dy_hi))."""
    _req(dy_hi, torch.bfloat16, "dy_hi")
    _req(wupT, torch.bfloat16, "wupT")
    n, d2, h2, w2, co = dy_hi.shape
    ci = wupT.shape[1]
    assert wupT.shape == (64, ci, co) and d2 % 2 == 0 and h2 % 2 == 0 and w2 % 2 == 0
    d, h, w = d2 // 2, h2 // 2, w2 // 2
    dx = torch.empty(n, d, h, w, ci, dtype=torch.bfloat16, device=dy_hi.device)
    flops = 2.0 * 27 * ci * co * n * d * h * w * 8
    _timed("upconv3_dgrad", (flops, (n, d, h, w, ci, co)),
           lambda: _check(_L().sivae_upconv3_dgrad(_p(dy_hi), _p(wupT), _p(dx), n, d, h, w, ci, co, _stream(dy_hi)),
                          "sivae_upconv3_dgrad"))
    return dx


def upconv3_wgrad(x_lo: torch.Tensor, dy_hi: torch.Tensor) -> torch.Tensor:
    """dw fp32 [Co,Ci,3,3,3] of conv3(upsample2(x_lo), w) given dy_hi."""
    _req(x_lo, torch.bfloat16, "x_lo")
    _req(dy_hi, torch.bfloat16, "dy_hi")
    n, d, h, w, ci = x_lo.shape
    co = dy_hi.shape[-1]
    assert tuple(dy_hi.shape[:4]) == (n, 2 * d, 2 * h, 2 * w)
    lib = _L()
    ws = _workspace(x_lo.device, lib.sivae_upconv3_wgrad_workspace_bytes(n, d, h, w, ci, co), "wgrad")
    dw = torch.empty(co, ci, 3, 3, 3, dtype=torch.float32, device=x_lo.device)
    flops = 2.0 * 27 * ci * co * n * d * h * w * 8
    _timed("upconv3_wgrad", (flops, (n, d, h, w, ci, co)),
           lambda: _check(lib.sivae_upconv3_wgrad(_p(x_lo), _p(dy_hi), _p(dw), _p(ws), ws.numel(), n, d, h, w, ci, co,
                                                  _stream(x_lo)), "sivae_upconv3_wgrad"))
    return dw


# ----------------------------------------------------------------------------------------------
# BatchNorm (train) + activation + residual + resample + dropout
# ----------------------------------------------------------------------------------------------
def bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum: float, eps: float):
    """-> (mean, invstd, scale, shift), each fp32 [C]; running stats are updated in place."""
    _req(y, torch.bfloat16, "y")
    c = y.shape[-1]
    nvox = y.numel() // c
    lib = _L()
    ws = _workspace(y.device, lib.sivae_bn_workspace_bytes(c), "bn")
    coef = torch.empty(4, c, dtype=torch.float32, device=y.device)
    if num_batches_tracked is not None:
        _req(num_batches_tracked, torch.int64, "num_batches_tracked")
    _check(lib.sivae_bn_train_coeffs(_p(y), nvox, c, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                     _p(num_batches_tracked), momentum, eps, _p(coef[0]), _p(coef[1]), _p(coef[2]),
                                     _p(coef[3]), _p(ws), ws.numel(), _stream(y)), "sivae_bn_train_coeffs")
    return coef[0], coef[1], coef[2], coef[3]


def _resampled_shape(shape, resample):
    n, d, h, w, c = shape
    if resample == RESAMPLE_AVGPOOL2:
        return (n, d // 2, h // 2, w // 2, c)
    if resample == RESAMPLE_UPSAMPLE2:
        return (n, 2 * d, 2 * h, 2 * w, c)
    return (n, d, h, w, c)


def _mask_args(y, mask, keep_bits, p, seed):
    """-> (pointer, seed) of the C ABI's dual-use ``mask`` argument: a caller-provided byte keep-mask travels with
    seed 0; the keep-bit store (uint8 [numel/8], written by the forward kernel from Philox(seed), read by backward)
    travels with the non-zero seed of its dropout call."""
    if mask is not None:
        _req(mask, torch.uint8, "mask")
        assert mask.shape == y.shape and keep_bits is None
        return _p(mask), 0
    if keep_bits is not None and p > 0.0:
        _req(keep_bits, torch.uint8, "keep_bits")
        assert keep_bits.numel() * 8 == y.numel() and seed != 0
        return _p(keep_bits), seed
    return None, seed


def bn_act_fwd(y, scale, shift, res, slope: float, resample: int, mask=None, p: float = 0.0, seed: int = 0,
               keep_bits=None):
    """out = resample(dropout(act(y*scale+shift (+res)))).  ``keep_bits`` (uint8 [numel/8], resample = none): the kernel
    stores its Philox keep decisions there, one bit per element, for ``bn_act_bwd``."""
    _req(y, torch.bfloat16, "y")
    n, d, h, w, c = y.shape
    if res is not None:
        _req(res, torch.bfloat16, "res")
        assert res.shape == y.shape
    mptr, seed = _mask_args(y, mask, keep_bits, p, seed)
    out = torch.empty(_resampled_shape(y.shape, resample), dtype=torch.bfloat16, device=y.device)
    _check(_L().sivae_bn_act_fwd(_p(y), _p(scale), _p(shift), _p(res), _p(out), n, d, h, w, c, slope, resample,
                                 mptr, p, seed, _stream(y)), "sivae_bn_act_fwd")
    return out


SMALL_BN_ELEMS = 3 << 20     # tensors up to this size take the single-launch cluster kernel of sivae_bn_train_act_fwd


def bn_train_act_fwd(y, res, gamma, beta, running_mean, running_var, num_batches_tracked, momentum: float, eps: float,
                     slope: float):
    """out = act(bn_train(y) (+ res)) with the batch statistics, coefficients and the apply in one C-ABI call
    (single launch for small tensors).  -> (out, mean, invstd)."""
    _req(y, torch.bfloat16, "y")
    n, d, h, w, c = y.shape
    if res is not None:
        _req(res, torch.bfloat16, "res")
        assert res.shape == y.shape
    if num_batches_tracked is not None:
        _req(num_batches_tracked, torch.int64, "num_batches_tracked")
    lib = _L()
    ws = _workspace(y.device, lib.sivae_bn_workspace_bytes(c), "bn")
    out = torch.empty_like(y)
    coef = torch.empty(4, c, dtype=torch.float32, device=y.device)
    _check(lib.sivae_bn_train_act_fwd(_p(y), _p(res), _p(out), n, d, h, w, c, _p(gamma), _p(beta), _p(running_mean),
                                      _p(running_var), _p(num_batches_tracked), momentum, eps, slope, _p(coef[0]),
                                      _p(coef[1]), _p(coef[2]), _p(coef[3]), _p(ws), ws.numel(), _stream(y)),
           "sivae_bn_train_act_fwd")
    return out, coef[0], coef[1]


def bn_act_bwd(g, y, res, mean, invstd, gamma, beta, slope: float, resample: int, mask=None, p: float = 0.0,
               seed: int = 0, need_dres: bool = False, need_affine: bool = True, keep_bits=None):
    """-> (dconv bf16 like y, dres bf16 or None, dgamma fp32 [C] or None, dbeta fp32 [C] or None)."""
    _req(g, torch.bfloat16, "g")
    _req(y, torch.bfloat16, "y")
    n, d, h, w, c = y.shape
    mptr, seed = _mask_args(y, mask, keep_bits, p, seed)
    assert tuple(g.shape) == _resampled_shape(y.shape, resample), (g.shape, y.shape, resample)
    lib = _L()
    ws = _workspace(y.device, lib.sivae_bn_workspace_bytes(c), "bn")
    dconv = torch.empty_like(y)
    dres = torch.empty_like(y) if need_dres else None
    aff = torch.empty(2, c, dtype=torch.float32, device=y.device) if need_affine else None
    _check(lib.sivae_bn_act_bwd(_p(g), _p(y), _p(res), _p(mean), _p(invstd), _p(gamma), _p(beta), _p(dconv), _p(dres),
                                _p(aff[0]) if need_affine else None, _p(aff[1]) if need_affine else None,
                                n, d, h, w, c, slope, resample, mptr, p, seed, _p(ws), ws.numel(), _stream(y)),
           "sivae_bn_act_bwd")
    return dconv, dres, (aff[0] if need_affine else None), (aff[1] if need_affine else None)


# ----------------------------------------------------------------------------------------------
# introspective loss assembly on the per-sample [B] vectors (utils/my_trainer.py:260-284, :301-321)
# ----------------------------------------------------------------------------------------------
def _vecs(*ts):
    b = ts[0].numel()
    for t in ts:
        _req(t, torch.float32, "per-sample vector")
        assert t.numel() == b
    return b


def intro_loss_e_fwd(r_real, k_real, r_fake, k_fake, r_rec, k_rec, scale, b_rec, b_kl, b_neg) -> torch.Tensor:
    """-> fp32 [5] = (lossE, mean(r_real), mean(k_real), exp_elbo_fake, exp_elbo_rec)."""
    b = _vecs(r_real, k_real, r_fake, k_fake, r_rec, k_rec)
    out = torch.empty(5, dtype=torch.float32, device=r_real.device)
    _check(_L().sivae_intro_loss_e_fwd(_p(r_real), _p(k_real), _p(r_fake), _p(k_fake), _p(r_rec), _p(k_rec), b, scale,
                                       b_rec, b_kl, b_neg, _p(out), _stream(r_real)), "sivae_intro_loss_e_fwd")
    return out


def intro_loss_e_bwd(r_fake, k_fake, r_rec, k_rec, g, scale, b_rec, b_kl, b_neg) -> torch.Tensor:
    """-> fp32 [6, B]: gradients w.r.t. (r_real, k_real, r_fake, k_fake, r_rec, k_rec) of lossE * g."""
    b = _vecs(r_fake, k_fake, r_rec, k_rec)
    _req(g, torch.float32, "g")
    d = torch.empty(6, b, dtype=torch.float32, device=r_fake.device)
    _check(_L().sivae_intro_loss_e_bwd(_p(r_fake), _p(k_fake), _p(r_rec), _p(k_rec), _p(g), b, scale, b_rec, b_kl, b_neg,
                                       _p(d[0]), _p(d[1]), _p(d[2]), _p(d[3]), _p(d[4]), _p(d[5]), _stream(r_fake)),
           "sivae_intro_loss_e_bwd")
    return d


def intro_loss_d_fwd(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, scale, b_rec, b_kl, gamma_r) -> torch.Tensor:
    """-> fp32 [6] = (lossD, mean(r_real), mean(k_rec), mean(k_fake), mean(r_rec_rec), mean(r_fake_rec))."""
    b = _vecs(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec)
    out = torch.empty(6, dtype=torch.float32, device=r_real.device)
    _check(_L().sivae_intro_loss_d_fwd(_p(r_real), _p(k_rec), _p(k_fake), _p(r_rec_rec), _p(r_fake_rec), b, scale, b_rec,
                                       b_kl, gamma_r, _p(out), _stream(r_real)), "sivae_intro_loss_d_fwd")
    return out


def intro_loss_d_bwd(g, batch: int, scale, b_rec, b_kl, gamma_r) -> torch.Tensor:
    """-> fp32 [5, B]: gradients w.r.t. (r_real, k_rec, k_fake, r_rec_rec, r_fake_rec) of lossD * g."""
    _req(g, torch.float32, "g")
    d = torch.empty(5, batch, dtype=torch.float32, device=g.device)
    _check(_L().sivae_intro_loss_d_bwd(_p(g), batch, scale, b_rec, b_kl, gamma_r, _p(d[0]), _p(d[1]), _p(d[2]), _p(d[3]),
                                       _p(d[4]), _stream(g)), "sivae_intro_loss_d_bwd")
    return d


# ----------------------------------------------------------------------------------------------
# fused multi-tensor Adam step + weight re-pack (SURVEY section 8f NEXT-2)
# ----------------------------------------------------------------------------------------------
class _AdamTensor(ctypes.Structure):
    """mirror of ``sivae_adam_tensor`` (include/sivae.h)"""
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("numel", _ll),
                ("pack_fwd", _vp), ("pack_dgrad", _vp), ("cout", _i), ("cin", _i)]


def adam_step(tensors, lr: torch.Tensor, beta1: float, beta2: float, eps: float, step: torch.Tensor):
    """One Adam update of every (param, grad, exp_avg, exp_avg_sq, packs) in ``tensors`` -- ``packs`` is None or the
    (wf, wd) bf16 packs of a Conv3d(k=3) weight, refreshed in the same pass.  ``lr`` (fp32) and ``step`` (int64) are
    one-element device tensors; ``step`` is incremented once."""
    if not tensors:
        return
    _req(lr, torch.float32, "lr")
    _req(step, torch.int64, "step")
    arr = (_AdamTensor * len(tensors))()
    for i, (p, g, m, v, packs) in enumerate(tensors):
        for t, nm in ((p, "param"), (g, "grad"), (m, "exp_avg"), (v, "exp_avg_sq")):
            _req(t, torch.float32, nm)
        assert g.shape == p.shape and m.shape == p.shape and v.shape == p.shape
        arr[i].param, arr[i].grad, arr[i].exp_avg, arr[i].exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
        arr[i].numel = p.numel()
        if packs is not None:
            wf, wd = packs
            assert p.dim() == 5 and tuple(p.shape[2:]) == (3, 3, 3)
            _req(wf, torch.bfloat16, "wf")
            arr[i].pack_fwd = wf.data_ptr()
            arr[i].pack_dgrad = wd.data_ptr() if wd is not None else None
            arr[i].cout, arr[i].cin = p.shape[0], p.shape[1]
    _check(_L().sivae_adam_step(ctypes.cast(arr, ctypes.c_void_p), len(tensors), _p(lr), beta1, beta2, eps, _p(step),
                                _stream(lr)), "sivae_adam_step")


# ----------------------------------------------------------------------------------------------
# on-GPU input pipeline (SURVEY section 8f NEXT-3)
# ----------------------------------------------------------------------------------------------
def volume_stats(x: torch.Tensor) -> torch.Tensor:
    """x fp32 [B, ...] -> [B, 4] = (mean, population std, min, max) per volume."""
    _req(x, torch.float32, "x")
    b = x.shape[0]
    n = x.numel() // b
    lib = _L()
    ws = _workspace(x.device, lib.sivae_volume_stats_workspace_bytes(b), "vstats")
    stats = torch.empty(b, 4, dtype=torch.float32, device=x.device)
    _check(lib.sivae_volume_stats(_p(x), b, n, _p(stats), _p(ws), ws.numel(), _stream(x)), "sivae_volume_stats")
    return stats


def preprocess_clip_minmax(x: torch.Tensor, cut_range: float = 4.0, out: Optional[torch.Tensor] = None):
    """BrainDataset._preprocess per volume: minmax(clip(x, 0, cut_range*std)).  -> (y, raw stats [B,4])."""
    _req(x, torch.float32, "x")
    b = x.shape[0]
    n = x.numel() // b
    lib = _L()
    ws = _workspace(x.device, lib.sivae_volume_stats_workspace_bytes(b), "vstats")
    stats = torch.empty(b, 4, dtype=torch.float32, device=x.device)
    y = torch.empty_like(x) if out is None else out
    _check(lib.sivae_preprocess_clip_minmax(_p(x), _p(y), b, n, float(cut_range), _p(stats), _p(ws), ws.numel(),
                                            _stream(x)), "sivae_preprocess_clip_minmax")
    return y, stats


def affine_resample(x: torch.Tensor, mats: torch.Tensor, pad: Optional[torch.Tensor] = None,
                    stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Trilinear resampling of x fp32 [B,D,H,W] through mats fp32 [B,12] (output voxel -> input voxel, rows d,h,w)."""
    _req(x, torch.float32, "x")
    _req(mats, torch.float32, "mats")
    b, d, h, w = x.shape
    assert mats.shape == (b, 12)
    y = torch.empty_like(x)
    _check(_L().sivae_affine_resample(_p(x), _p(y), b, d, h, w, _p(mats), _p(pad), _p(stats), _stream(x)),
           "sivae_affine_resample")
    return y


# ----------------------------------------------------------------------------------------------
# latent retrieval (SURVEY section 8f NEXT-4)
# ----------------------------------------------------------------------------------------------
def similarity_topk(queries: torch.Tensor, database: torch.Tensor, k: int, metric: str = "cosine"):
    """-> (scores [nq,k] fp32, index [nq,k] int32), best first.  metric "cosine" or "l2" (score = -squared distance)."""
    _req(queries, torch.float32, "queries")
    _req(database, torch.float32, "database")
    nq, dim = queries.shape
    nd, dim2 = database.shape
    assert dim == dim2
    if metric not in ("cosine", "l2"):
        raise ValueError("metric must be 'cosine' or 'l2'")
    lib = _L()
    ws = _workspace(queries.device, lib.sivae_similarity_workspace_bytes(nq, nd), "sim")
    scores = torch.empty(nq, k, dtype=torch.float32, device=queries.device)
    index = torch.empty(nq, k, dtype=torch.int32, device=queries.device)
    _check(lib.sivae_similarity_topk(_p(queries), _p(database), nq, nd, dim, 0 if metric == "cosine" else 1, k,
                                     _p(scores), _p(index), _p(ws), ws.numel(), _stream(queries)),
           "sivae_similarity_topk")
    return scores, index


# ----------------------------------------------------------------------------------------------
# FC-latent variant (SURVEY section 8f NEXT-1): Linear heads, layout changes, add + activation
# ----------------------------------------------------------------------------------------------
def linear_fwd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False):
    """y = act(x @ weight.T + bias): x fp32 [B,K], weight fp32 [J,K] (nn.Linear layout), bias fp32 [J] -> fp32 [B,J]."""
    _req(x, torch.float32, "x")
    _req(weight, torch.float32, "weight")
    b, k = x.shape
    j, k2 = weight.shape
    assert k == k2 and (bias is None or bias.numel() == j)
    lib = _L()
    ws = _workspace(x.device, lib.sivae_linear_workspace_bytes(b, k, j), "linear")
    y = torch.empty(b, j, dtype=torch.float32, device=x.device)
    _check(lib.sivae_linear_fwd(_p(x), _p(weight), _p(bias), _p(y), b, k, j, int(relu), _p(ws), ws.numel(), _stream(x)),
           "sivae_linear_fwd")
    return y


def linear_dgrad(dy: torch.Tensor, weight: torch.Tensor):
    """dx = dy @ weight: dy fp32 [B,J], weight fp32 [J,K] -> fp32 [B,K]."""
    _req(dy, torch.float32, "dy")
    _req(weight, torch.float32, "weight")
    b, j = dy.shape
    j2, k = weight.shape
    assert j == j2
    lib = _L()
    ws = _workspace(dy.device, lib.sivae_linear_workspace_bytes(b, k, j), "linear")
    dx = torch.empty(b, k, dtype=torch.float32, device=dy.device)
    _check(lib.sivae_linear_dgrad(_p(dy), _p(weight), _p(dx), b, k, j, _p(ws), ws.numel(), _stream(dy)),
           "sivae_linear_dgrad")
    return dx


def linear_wgrad(x: torch.Tensor, dy: torch.Tensor, need_bias: bool = True):
    """-> (dW fp32 [J,K] = dy.T @ x, db fp32 [J] = dy.sum(0) or None)."""
    _req(x, torch.float32, "x")
    _req(dy, torch.float32, "dy")
    b, k = x.shape
    b2, j = dy.shape
    assert b == b2
    dw = torch.empty(j, k, dtype=torch.float32, device=x.device)
    db = torch.empty(j, dtype=torch.float32, device=x.device) if need_bias else None
    _check(_L().sivae_linear_wgrad(_p(x), _p(dy), _p(dw), _p(db), b, k, j, _stream(x)), "sivae_linear_wgrad")
    return dw, db


def ndhwc_to_flat(h: torch.Tensor, c: int, gate: Optional[torch.Tensor] = None):
    """NDHWC bf16 [B,d,h,w,Cp] -> fp32 [B, c*d*h*w] in NCDHW flatten order (x.view(B,-1), mymodel.py:140); entries whose
    ``gate`` (fp32, same shape as the result) is <= 0 are zeroed."""
    _req(h, torch.bfloat16, "h")
    b, cp = h.shape[0], h.shape[-1]
    s = h.numel() // (b * cp)
    out = torch.empty(b, c * s, dtype=torch.float32, device=h.device)
    if gate is not None:
        _req(gate, torch.float32, "gate")
        assert gate.shape == out.shape
    _check(_L().sivae_ndhwc_to_flat(_p(h), _p(out), b, s, c, cp, _p(gate), _stream(h)), "sivae_ndhwc_to_flat")
    return out


def flat_to_ndhwc(y: torch.Tensor, c: int, cp: int, grid):
    """fp32 [B, c*S] (NCDHW flatten order, S = prod(grid)) -> NDHWC bf16 [B,*grid,cp], channels zero-padded."""
    _req(y, torch.float32, "y")
    b = y.shape[0]
    s = grid[0] * grid[1] * grid[2]
    assert y.shape[1] == c * s
    out = torch.empty(b, grid[0], grid[1], grid[2], cp, dtype=torch.bfloat16, device=y.device)
    _check(_L().sivae_flat_to_ndhwc(_p(y), _p(out), b, s, c, cp, _stream(y)), "sivae_flat_to_ndhwc")
    return out


def add_act_fwd(a: torch.Tensor, b: torch.Tensor, slope: float):
    """LeakyReLU_slope(a + b) on bf16 tensors of equal shape (mymodel.py:136)."""
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.shape == b.shape
    out = torch.empty_like(a)
    _check(_L().sivae_add_act_fwd(_p(a), _p(b), _p(out), a.numel(), float(slope), _stream(a)), "sivae_add_act_fwd")
    return out


def add_act_bwd(g: torch.Tensor, out: torch.Tensor, slope: float):
    """Gradient of add_act_fwd with respect to either operand: g * (out > 0 ? 1 : slope)."""
    _req(g, torch.bfloat16, "g")
    _req(out, torch.bfloat16, "out")
    dz = torch.empty_like(g)
    _check(_L().sivae_add_act_bwd(_p(g), _p(out), _p(dz), g.numel(), float(slope), _stream(g)), "sivae_add_act_bwd")
    return dz


# ----------------------------------------------------------------------------------------------
# thin convolutions (one channel on one side)
# ----------------------------------------------------------------------------------------------
def c1_to_cn(x1, w, bias, flip: bool = False, out: Optional[torch.Tensor] = None):
    """y[v,c] (+)= bias[c] + sum_t w[c,t]*x1[v+delta(t)].  x1 fp32 [N,D,H,W]; w fp32 [C,T]; out given => accumulate."""
    _req(x1, torch.float32, "x1")
    _req(w, torch.float32, "w")
    n, d, h, ww = x1.shape
    c, t = w.shape
    acc = out is not None
    if out is None:
        out = torch.empty(n, d, h, ww, c, dtype=torch.bfloat16, device=x1.device)
    else:
        _req(out, torch.bfloat16, "out")
    if t == 27 and c == 64 and not acc:
        # tensor-core path: im2col built in shared memory, hi/lo bf16 split of both operands
        lib = _L()
        ws = _workspace(x1.device, lib.sivae_c1_to_c64_workspace_bytes(), "c1c64")
        _check(lib.sivae_c1_to_c64(_p(x1), _p(w), _p(bias), _p(out), n, d, h, ww, int(flip), _p(ws), ws.numel(),
                                   _stream(x1)), "sivae_c1_to_c64")
        return out
    _check(_L().sivae_c1_to_cn(_p(x1), _p(w), _p(bias), _p(out), n, d, h, ww, c, t, int(flip), int(acc), _stream(x1)),
           "sivae_c1_to_cn")
    return out


def c1_to_cn_bn(x1, w, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum: float, eps: float):
    """Stem convolution + train-mode BatchNorm coefficients of its output: -> (y, mean, invstd, scale, shift).
    Conv3d(1,64,3) runs as one fused call (channel sums from the convolution epilogue); other widths / 1x1 kernels
    run the convolution and the statistics pass separately."""
    c, t = w.shape
    if not (t == 27 and c == 64):
        y = c1_to_cn(x1, w, bias)
        return (y,) + tuple(bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps))
    _req(x1, torch.float32, "x1")
    _req(w, torch.float32, "w")
    n, d, h, ww = x1.shape
    lib = _L()
    y = torch.empty(n, d, h, ww, c, dtype=torch.bfloat16, device=x1.device)
    ws = _workspace(x1.device, lib.sivae_c1_to_c64_workspace_bytes(), "c1c64")
    wsb = _workspace(x1.device, lib.sivae_bn_workspace_bytes(c), "bn")
    coef = torch.empty(4, c, dtype=torch.float32, device=x1.device)
    if num_batches_tracked is not None:
        _req(num_batches_tracked, torch.int64, "num_batches_tracked")
    _check(lib.sivae_c1_to_c64_bn(_p(x1), _p(w), _p(bias), _p(y), n, d, h, ww, 0, _p(gamma), _p(beta), _p(running_mean),
                                  _p(running_var), _p(num_batches_tracked), momentum, eps, _p(coef[0]), _p(coef[1]),
                                  _p(coef[2]), _p(coef[3]), _p(ws), ws.numel(), _p(wsb), wsb.numel(), _stream(x1)),
           "sivae_c1_to_c64_bn")
    return y, coef[0], coef[1], coef[2], coef[3]


def cn_to_c1(x, w, bias, flip: bool = False, act: int = 0, mask=None, p: float = 0.0, seed: int = 0):
    """y[v] = act(bias + sum_{t,c} w[c,t]*x[v+delta(t),c]).  x bf16 NDHWC; w fp32 [C,T]; y fp32 [N,D,H,W]."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.float32, "w")
    n, d, h, ww, c = x.shape
    assert w.shape[0] == c
    if mask is not None:
        _req(mask, torch.uint8, "mask")
    y = torch.empty(n, d, h, ww, dtype=torch.float32, device=x.device)
    lib = _L()
    if w.shape[1] == 27 and c % 64 == 0:
        # 3x3x3: implicit GEMM on tcgen05 (N=16 tile, fp32 column-0 epilogue); weights are rounded to bf16
        ws = _workspace(x.device, lib.sivae_conv3_to1_workspace_bytes(c), "to1")
        flops = 2.0 * 27 * c * 1 * n * d * h * ww          # reference-equivalent: Cout = 1 (the kernel pads N to 16 / 64 columns)
        _timed("conv3_to1", (flops, (n, d, h, ww, c, 1)),
               lambda: _check(lib.sivae_conv3_to1(_p(x), _p(w), _p(bias), _p(y), n, d, h, ww, c, int(flip), act,
                                                  _p(mask), p, seed, _p(ws), ws.numel(), _stream(x)),
                              "sivae_conv3_to1"))
        return y
    _check(lib.sivae_cn_to_c1(_p(x), _p(w), _p(bias), _p(y), n, d, h, ww, c, w.shape[1], int(flip), act, _p(mask), p,
                              seed, _stream(x)), "sivae_cn_to_c1")
    return y


def wgrad_c1(xc, x1, taps: int, flip: bool = False):
    """-> (dw fp32 [C,T], sum_c fp32 [C] = sum_v xc[v,c], sum_1 fp32 [1] = sum_v x1[v])."""
    _req(xc, torch.bfloat16, "xc")
    _req(x1, torch.float32, "x1")
    n, d, h, ww, c = xc.shape
    lib = _L()
    dw = torch.empty(c, taps, dtype=torch.float32, device=xc.device)
    sum_c = torch.empty(c, dtype=torch.float32, device=xc.device)
    sum_1 = torch.empty(1, dtype=torch.float32, device=xc.device)
    if taps == 27 and c == 64:
        ws = _workspace(xc.device, lib.sivae_wgrad_c64_workspace_bytes(), "wgrad_c64")
        _check(lib.sivae_wgrad_c64(_p(xc), _p(x1), _p(dw), _p(sum_c), _p(sum_1), n, d, h, ww, int(flip), _p(ws),
                                   ws.numel(), _stream(xc)), "sivae_wgrad_c64")
        return dw, sum_c, sum_1
    ws = _workspace(xc.device, lib.sivae_wgrad_c1_workspace_bytes(n, d, h, ww, c, taps), "wgrad_c1")
    _check(lib.sivae_wgrad_c1(_p(xc), _p(x1), _p(dw), _p(sum_c), _p(sum_1), n, d, h, ww, c, taps, int(flip), _p(ws),
                              ws.numel(), _stream(xc)), "sivae_wgrad_c1")
    return dw, sum_c, sum_1


def relu_drop_bwd(g, out, p: float):
    _req(g, torch.float32, "g")
    _req(out, torch.float32, "out")
    dy = torch.empty_like(out)
    _check(_L().sivae_relu_drop_bwd(_p(g), _p(out), _p(dy), out.numel(), p, _stream(out)), "sivae_relu_drop_bwd")
    return dy


# ----------------------------------------------------------------------------------------------
# latent / loss
# ----------------------------------------------------------------------------------------------
def reparam_fwd(mu, logvar, eps):
    """eps: fp32 tensor like mu, or a python float (validation: 0.1)."""
    _req(mu, torch.float32, "mu")
    _req(logvar, torch.float32, "logvar")
    z = torch.empty_like(mu)
    et = eps if isinstance(eps, torch.Tensor) else None
    if et is not None:
        _req(et, torch.float32, "eps")
    _check(_L().sivae_reparam_fwd(_p(mu), _p(logvar), _p(et), 0.0 if et is not None else float(eps), _p(z),
                                  mu.numel(), _stream(mu)), "sivae_reparam_fwd")
    return z


def reparam_draw_fwd(mu, logvar, seed: int):
    """Training-path sampler: eps ~ N(0,1) drawn in the kernel (Philox keyed by ``seed`` and the device epoch counter).
    -> (z, eps)."""
    _req(mu, torch.float32, "mu")
    _req(logvar, torch.float32, "logvar")
    assert mu.shape == logvar.shape
    z, eps = torch.empty_like(mu), torch.empty_like(mu)
    _check(_L().sivae_reparam_draw_fwd(_p(mu), _p(logvar), _p(eps), _p(z), mu.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                       _stream(mu)), "sivae_reparam_draw_fwd")
    return z, eps


def reparam_bwd(dz, logvar, eps):
    _req(dz, torch.float32, "dz")
    dmu, dlv = torch.empty_like(dz), torch.empty_like(dz)
    et = eps if isinstance(eps, torch.Tensor) else None
    _check(_L().sivae_reparam_bwd(_p(dz), _p(logvar), _p(et), 0.0 if et is not None else float(eps), _p(dmu), _p(dlv),
                                  dz.numel(), 0, _stream(dz)), "sivae_reparam_bwd")
    return dmu, dlv


def kl_persample_fwd(mu, logvar):
    """mu, logvar: fp32 [B, n] -> kl fp32 [B]."""
    _req(mu, torch.float32, "mu")
    _req(logvar, torch.float32, "logvar")
    b, n = mu.shape
    kl = torch.empty(b, dtype=torch.float32, device=mu.device)
    _check(_L().sivae_kl_persample_fwd(_p(mu), _p(logvar), _p(kl), b, n, _stream(mu)), "sivae_kl_persample_fwd")
    return kl


def kl_persample_bwd(mu, logvar, g):
    _req(g, torch.float32, "g")
    b, n = mu.shape
    dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
    _check(_L().sivae_kl_persample_bwd(_p(mu), _p(logvar), _p(g), _p(dmu), _p(dlv), b, n, 0, _stream(mu)),
           "sivae_kl_persample_bwd")
    return dmu, dlv


def mse_persample_fwd(x, y):
    """x, y: fp32 [B, n] -> r fp32 [B], r[b] = sum_j (x-y)^2."""
    _req(x, torch.float32, "x")
    _req(y, torch.float32, "y")
    b, n = x.shape
    lib = _L()
    ws = _workspace(x.device, lib.sivae_mse_workspace_bytes(b, n), "mse")
    r = torch.empty(b, dtype=torch.float32, device=x.device)
    _check(lib.sivae_mse_persample_fwd(_p(x), _p(y), _p(r), b, n, _p(ws), ws.numel(), _stream(x)),
           "sivae_mse_persample_fwd")
    return r


def mse_persample_bwd(x, y, g, need_dx: bool, need_dy: bool):
    _req(g, torch.float32, "g")
    b, n = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dy = torch.empty_like(y) if need_dy else None
    _check(_L().sivae_mse_persample_bwd(_p(x), _p(y), _p(g), _p(dx), _p(dy), b, n, _stream(x)),
           "sivae_mse_persample_bwd")
    return dx, dy


# ----------------------------------------------------------------------------------------------
# layout helpers
# ----------------------------------------------------------------------------------------------
def to_ndhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCDHW [N,C,D,H,W] -> bf16 NDHWC [N,D,H,W,C]."""
    _req(x, torch.float32, "x")
    n, c, d, h, w = x.shape
    out = torch.empty(n, d, h, w, c, dtype=torch.bfloat16, device=x.device)
    _check(_L().sivae_ncdhw_f32_to_ndhwc_bf16(_p(x), _p(out), n, c, d * h * w, _stream(x)), "sivae_ncdhw_f32_to_ndhwc_bf16")
    return out


def to_ncdhw_f32(x: torch.Tensor) -> torch.Tensor:
    """bf16 NDHWC [N,D,H,W,C] -> fp32 NCDHW [N,C,D,H,W]."""
    _req(x, torch.bfloat16, "x")
    n, d, h, w, c = x.shape
    out = torch.empty(n, c, d, h, w, dtype=torch.float32, device=x.device)
    _check(_L().sivae_ndhwc_bf16_to_ncdhw_f32(_p(x), _p(out), n, c, d * h * w, _stream(x)), "sivae_ndhwc_bf16_to_ncdhw_f32")
    return out
