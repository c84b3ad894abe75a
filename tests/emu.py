"""CPU emulation of libsivae.so for wiring tests: monkeypatches every function of
``sivae_b200.kernels`` with its executable specification from ``oracle/kernel_spec.py``.
Test infrastructure only -- the product never imports this."""
import contextlib

import torch

import sivae_b200
from oracle import kernel_spec as S


@contextlib.contextmanager
def emulated_kernels(act_dtype=torch.float32):
    K = sivae_b200.kernels
    saved = {}
    old_dtype = S.ACT_DTYPE
    S.ACT_DTYPE = act_dtype
    names = [n for n in S.ALL if hasattr(K, n) and callable(getattr(K, n))]
    for n in names:
        saved[n] = getattr(K, n)
        setattr(K, n, getattr(S, n))
    sivae_b200.functional._pack_cache.clear()
    try:
        yield names
    finally:
        for n, f in saved.items():
            setattr(K, n, f)
        S.ACT_DTYPE = old_dtype
        sivae_b200.functional._pack_cache.clear()


def masks_to_feed(masks, cpad=64):
    """golden NCDHW bool keep-masks -> the uint8 NDHWC (channel-padded) / [N,D,H,W] masks the kernels take."""
    out = []
    for m in masks:
        if m.shape[1] == 1:
            out.append(m[:, 0].to(torch.uint8).contiguous())
        else:
            c = m.shape[1]
            cp = (c + cpad - 1) // cpad * cpad
            mm = m.permute(0, 2, 3, 4, 1).to(torch.uint8)
            if cp != c:
                mm = torch.cat([mm, torch.ones(*mm.shape[:-1], cp - c, dtype=torch.uint8)], dim=-1)
            out.append(mm.contiguous())
    return out
