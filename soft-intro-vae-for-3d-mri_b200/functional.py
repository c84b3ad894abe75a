"""Autograd wiring of the fused units on top of ``kernels`` (one C-ABI call per line).

Units (all activations NDHWC, one-channel tensors fp32 [N,D,H,W]):
  conv_bn_act       Conv3d(3, bias=False) -> BatchNorm3d -> (Leaky)ReLU [+res] -> AvgPool/Upsample
                    (reference: BuildingBlock / UpsampleBuildingkBlock, models/models.py:8-80)
  stem_bn_act       Conv3d(1, C, k, bias) -> BatchNorm3d -> (Leaky)ReLU -> Dropout
                    (encoder stem models.py:91-96, k=3; decoder stem :117-123, k=1)
  tail_relu_drop    Conv3d(C, 1, 3, bias) -> ReLU -> Dropout          (models.py:136-141)
  heads             mu / var 1x1 Conv3d(C, 1)                          (models.py:216-223)
  reparameterize, kl_persample, mse_persample                          (models.py:263-271,
                    utils/my_trainer.py:38-78)
Backward follows SURVEY.md appendix B; parameter gradients are only computed when the parameter
requires grad (the trainer freezes encoder / decoder per phase, utils/my_trainer.py:242-245,291-294).
"""
from __future__ import annotations

import contextlib
import os
import weakref
from typing import Optional

import torch

from . import kernels as K

KEEP_BITS = os.environ.get("SIVAE_KEEP_BITS", "1") != "0"   # 0: regenerate Philox in both backward passes (A/B switch)
BN_MOMENTUM = 0.1
BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# dropout state: every dropout call gets a fresh 64-bit Philox key; backward reuses the same key
# ----------------------------------------------------------------------------------------------
class _DropoutState:
    """Philox keys of the in-kernel dropout.  Each dropout call of a training step gets the key
    (base_seed, call index within the step); a device-resident epoch counter, advanced once per step by
    ``begin_step``, is mixed in by the kernels, so the host-side constants can be baked into a CUDA graph
    and every replay still draws fresh masks.  Backward reuses the key saved by its forward."""

    def __init__(self):
        self.base_seed = 0x5EED5EED
        self.call = 0
        self.epoch = None       # int64 device tensor registered with libsivae (one device per process)
        self.mask_feed = None   # optional iterator of explicit uint8 keep-masks (parity tests)

    def next(self):
        """-> (mask or None, seed)"""
        self.call += 1
        seed = (self.base_seed * 0x9E3779B97F4A7C15 + self.call * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        mask = next(self.mask_feed) if self.mask_feed is not None else None
        return mask, seed

    def begin_step(self, device):
        """Call once at the start of every training step (graph-capturable)."""
        self.call = 0
        noise_state.call = 0
        device = torch.device(device)
        if device.type != "cuda":
            return
        if self.epoch is None or self.epoch.device != device:
            self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
            K.set_seed_counter(self.epoch)
        K.advance_seed_counter(device)


dropout_state = _DropoutState()


def manual_seed(seed: int):
    """Re-key the in-kernel Philox dropout streams (independent of torch's generator)."""
    dropout_state.base_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    dropout_state.call = 0
    noise_state.base_seed = (int(seed) * 0x9E3779B97F4A7C15 + 0x0E95) & 0xFFFFFFFFFFFFFFFF
    noise_state.call = 0
    if dropout_state.epoch is not None:
        dropout_state.epoch.zero_()


def begin_step(device):
    dropout_state.begin_step(device)


class BnState:
    """Non-differentiable BatchNorm state handed to the fused units.  ``defer`` = the holder's (running_mean,
    running_var, num_batches_tracked) when this call must NOT touch them itself (it runs on a side stream next to
    another pass through the same layer, see ``deferred_bn``): the kernels then get NULL buffers and the batch
    statistics are logged for ``apply_deferred_bn``."""

    __slots__ = ("running_mean", "running_var", "num_batches_tracked", "training", "momentum", "eps", "defer")

    def __init__(self, running_mean, running_var, num_batches_tracked, training, momentum=BN_MOMENTUM, eps=BN_EPS,
                 defer=None):
        self.running_mean, self.running_var, self.num_batches_tracked = running_mean, running_var, num_batches_tracked
        self.training, self.momentum, self.eps, self.defer = training, momentum, eps, defer


# ----------------------------------------------------------------------------------------------
# two independent passes on two streams (trainer._fork_join): the side branch defers its BatchNorm running-statistic
# updates, so that two passes through the same layer never race on the buffers and the updates compose in the
# reference's call order (running <- 0.9 running + 0.1 batch is order dependent)
# ----------------------------------------------------------------------------------------------
class _BnDefer:
    def __init__(self):
        self.records = None      # list while a side branch is being issued, else None


bn_defer = _BnDefer()


@contextlib.contextmanager
def deferred_bn():
    """Inside: train-mode BatchNorm calls leave the running statistics alone and log (buffers, mean, invstd, n, momentum,
    eps) instead.  -> the log, to be passed to ``apply_deferred_bn`` after the streams have joined."""
    prev, bn_defer.records = bn_defer.records, []
    try:
        yield bn_defer.records
    finally:
        bn_defer.records = prev


def _note_bn(bn: BnState, mean, invstd, nvox: int):
    if bn.defer is not None and bn.training:
        c = bn.defer[0].numel()                         # the holder's channel count (the kernels may run zero-padded)
        if mean.numel() != c:
            mean, invstd = mean[:c], invstd[:c]
        bn_defer.records.append([bn.defer, mean, invstd, int(nvox), float(bn.momentum), float(bn.eps), None])


def note_bias_into_running_mean(bias):
    """FC-latent variant (mymodel.py): a convolution bias in front of a train-mode BatchNorm only shifts the batch mean;
    the fused units run the bias-free convolution, so the deferred running-mean update of the record just logged gets
    ``+ momentum * bias`` (the non-deferred path adds it through mymodel._BiasIntoRunningMean)."""
    bn_defer.records[-1][6] = bias.detach()


def apply_deferred_bn(records):
    """running_mean <- (1-m) running_mean + m mean;  running_var <- (1-m) running_var + m var n/(n-1) with the biased batch
    variance var = 1/invstd^2 - eps;  num_batches_tracked += 1 -- in log order, a handful of multi-tensor launches per
    wave (wave k = the k-th pass through each layer in the log)."""
    if not records:
        return
    waves, count = [], {}
    for r in records:
        k = count.get(id(r[0][0]), 0)
        count[id(r[0][0])] = k + 1
        if k == len(waves):
            waves.append([])
        waves[k].append(r)
    with torch.no_grad():
        for wave in waves:
            for m in sorted({r[4] for r in wave}):
                rs = [r for r in wave if r[4] == m]
                var = torch._foreach_pow([r[2] for r in rs], -2.0)
                torch._foreach_sub_(var, [r[5] for r in rs])
                torch._foreach_mul_(var, [m * (r[3] / (r[3] - 1.0) if r[3] > 1 else 1.0) for r in rs])
                rms, rvs = [r[0][0] for r in rs], [r[0][1] for r in rs]
                torch._foreach_mul_(rms, 1.0 - m)
                torch._foreach_add_(rms, [r[1] for r in rs], alpha=m)
                with_bias = [r for r in rs if r[6] is not None]
                if with_bias:
                    torch._foreach_add_([r[0][0] for r in with_bias], [r[6] for r in with_bias], alpha=m)
                torch._foreach_mul_(rvs, 1.0 - m)
                torch._foreach_add_(rvs, var)
            nbts = [r[0][2] for r in wave if r[0][2] is not None]
            if nbts:
                torch._foreach_add_(nbts, 1)


# ----------------------------------------------------------------------------------------------
# weight gradients on their own stream (default on; SIVAE_WGRAD_STREAM=0 disables; 64.2 -> 63.6 ms alone, A/B on one box)
# ----------------------------------------------------------------------------------------------
class _WgradSideStream:
    """Backward's critical path is the chain  BN-backward -> dgrad -> BN-backward -> ...  (memory-bound passes alternating
    with tensor-bound ones); the weight gradient of a layer depends only on that layer's ``dconv`` and nothing waits
    for it until the optimiser.  Inside ``with wgrad_side_stream(device):`` the 3x3x3 weight-gradient kernels are
    therefore issued on a second stream, where they overlap the BatchNorm passes of the layers further down, and their
    results are accumulated THERE (several passes share a weight) and handed to ``param.grad`` when the block exits,
    after the streams have joined.  autograd sees no gradient for those weights (``None``), so this path is not used
    together with gradient hooks (parallel.GradReducer); parallel.FlatGradReducer's persistent ``.grad`` views work."""

    def __init__(self):
        self.active = False
        self.streams = {}
        self.stream = None
        self.pending = {}       # id(param) -> [param, fp32 sum of its weight gradients (lives on the side stream)]
        self.keep = []          # operands of in-flight kernels: freed only after the join

    def launch(self, params, fn, operands):
        """``fn()`` -> one gradient per entry of ``params`` (a Parameter, or a tuple of them), computed on the side stream."""
        single = not isinstance(params, (tuple, list))
        plist = [params] if single else list(params)
        cur = torch.cuda.current_stream(plist[0].device)
        self.stream.wait_stream(cur)                    # dconv is complete
        with torch.cuda.stream(self.stream):
            out = fn()
            for param, dw in zip(plist, [out] if single else out):
                if param is None or dw is None:
                    continue
                ent = self.pending.get(id(param))
                if ent is None:
                    self.pending[id(param)] = [param, dw]
                else:
                    ent[1].add_(dw.reshape(ent[1].shape))
        self.keep.append(operands)

    def flush(self, device):
        if not self.pending:
            return
        torch.cuda.current_stream(device).wait_stream(self.stream)
        with torch.no_grad():
            for param, buf in self.pending.values():
                buf = buf.reshape(param.shape)
                if param.grad is None:
                    param.grad = buf
                else:
                    param.grad.add_(buf)
        self.pending, self.keep = {}, []


def _owning_param(t):
    """The nn.Parameter ``t`` is, or is a plain view of (``conv.weight.reshape(C, 27)``); None for padded copies."""
    if isinstance(t, torch.nn.Parameter):
        return t
    base = getattr(t, "_base", None)
    if isinstance(base, torch.nn.Parameter) and base.numel() == t.numel():
        return base
    return None


wgrad_side = _WgradSideStream()
WGRAD_STREAM = os.environ.get("SIVAE_WGRAD_STREAM", "1") != "0"


@contextlib.contextmanager
def wgrad_side_stream(device, enabled=True):
    """Wrap ``loss.backward()``: 3x3x3 weight gradients of parameters passed unpadded run on a side stream."""
    device = torch.device(device)
    if not (enabled and device.type == "cuda"):
        yield
        return
    st = wgrad_side.streams.get(device)
    if st is None:
        st = wgrad_side.streams[device] = torch.cuda.Stream(device)
    wgrad_side.stream, wgrad_side.active = st, True
    try:
        yield
    finally:
        wgrad_side.active = False
        wgrad_side.flush(device)


def _bn_coeffs(y, gamma, beta, bn: BnState):
    """-> (mean, invstd, scale, shift).  Train: batch statistics (+ running-stat update).  Eval: running stats."""
    if bn.training:
        coef = K.bn_train_coeffs(y, gamma, beta, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                 bn.momentum, bn.eps)
        _note_bn(bn, coef[0], coef[1], y.numel() // y.shape[-1])
        return coef
    invstd = torch.rsqrt(bn.running_var + bn.eps)
    scale = gamma * invstd
    return bn.running_mean, invstd, scale, beta - bn.running_mean * scale


# Weight packs are cached per live tensor object and validated by (storage pointer, version): optimiser
# steps bump ``_version``; the trainer's per-epoch ``model.to('cpu')`` / ``.to(device)`` round trip
# re-allocates storage; temporaries (channel-padded weights) die and fail the weakref test.
_pack_cache = {}


def _packed(weight: torch.Tensor, up: bool = False):
    """-> (forward pack, data-gradient pack): plain 27-tap packs, or the upsample-folded 64-slab packs."""
    key = (id(weight), up)
    hit = _pack_cache.get(key)
    if hit is not None:
        ref, ptr, ver, packs = hit
        if ref() is weight and ptr == weight.data_ptr() and ver == weight._version:
            return packs
    if len(_pack_cache) > 512:
        for k in [k for k, v in _pack_cache.items() if v[0]() is None]:
            del _pack_cache[k]
    w = weight.detach().contiguous()
    packs = K.pack_upconv3_weights(w) if up else K.pack_conv3_weights(w)
    _pack_cache[key] = (weakref.ref(weight), weight.data_ptr(), weight._version, packs)
    return packs


def current_packs(weight: torch.Tensor, up: bool = False):
    """The cached packs of ``weight`` if they are current, else None (used by optim.FusedAdam, which refreshes the
    plain 27-tap packs inside its update kernel)."""
    hit = _pack_cache.get((id(weight), up))
    if hit is not None:
        ref, ptr, ver, packs = hit
        if ref() is weight and ptr == weight.data_ptr() and ver == weight._version:
            return packs
    return None


def drop_packs(weight: torch.Tensor, up: Optional[bool] = None):
    """Forget cached packs of ``weight`` (both kinds, or only the plain / upsample-folded one): an update that bypasses
    torch's version counter (FusedAdam) must call this for every pack it did not refresh itself."""
    for kind in ((False, True) if up is None else (up,)):
        _pack_cache.pop((id(weight), kind), None)


class _ConvBnAct(torch.autograd.Function):
    """``pre_up``: a nearest Upsample(2) sits in front of the convolution (UpsampleBuildingkBlock, models.py:58-59);
    it is folded into the convolution (8 output parities x 8 pre-summed taps) instead of being materialised."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, res, bn: BnState, slope: float, resample: int, pre_up: bool):
        wf, wd = _packed(weight, pre_up)
        small = (bn.training and resample == K.RESAMPLE_NONE
                 and x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] * (8 if pre_up else 1) * wf.shape[1] <= K.SMALL_BN_ELEMS)
        if small:
            # small layers (latent resolution): convolution, then statistics + coefficients + apply behind one C-ABI call
            # (a single cluster launch when SIVAE_BN_CLUSTER=1; three short launches by default, which measured faster)
            y = K.upconv3_fprop(x, wf) if pre_up else K.conv3_igemm(x, wf)
            out, mean, invstd = K.bn_train_act_fwd(y, res, gamma, beta, bn.running_mean, bn.running_var,
                                                   bn.num_batches_tracked, bn.momentum, bn.eps, slope)
            _note_bn(bn, mean, invstd, y.numel() // y.shape[-1])
            ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, y, res, mean, invstd, gamma, beta, wd)
            ctx.cfg = (slope, resample, bn.training, pre_up)
            ctx.wparam = weight if isinstance(weight, torch.nn.Parameter) else None
            return out
        if bn.training:
            # convolution + batch statistics in one C-ABI call (the sums come out of the conv epilogue)
            conv_bn = K.upconv3_fprop_bn if pre_up else K.conv3_igemm_bn
            y, mean, invstd, scale, shift = conv_bn(x, wf, gamma, beta, bn.running_mean, bn.running_var,
                                                    bn.num_batches_tracked, bn.momentum, bn.eps)
            _note_bn(bn, mean, invstd, y.numel() // y.shape[-1])
        else:
            y = K.upconv3_fprop(x, wf) if pre_up else K.conv3_igemm(x, wf)
            mean, invstd, scale, shift = _bn_coeffs(y, gamma, beta, bn)
        out = K.bn_act_fwd(y, scale, shift, res, slope, resample)
        # x is only needed for the weight gradient: frozen-parameter passes (dgrad-only) do not keep it
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, y, res, mean, invstd, gamma, beta, wd)
        ctx.cfg = (slope, resample, bn.training, pre_up)
        ctx.wparam = weight if isinstance(weight, torch.nn.Parameter) else None
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, res, mean, invstd, gamma, beta, wd = ctx.saved_tensors
        slope, resample, training, pre_up = ctx.cfg
        if not training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the reference hot path")
        need_x, need_w, need_g, need_b, need_res = ctx.needs_input_grad[:5]
        dconv, dres, dgamma, dbeta = K.bn_act_bwd(g.contiguous(), y, res, mean, invstd, gamma, beta, slope, resample,
                                                  need_dres=bool(need_res and res is not None),
                                                  need_affine=bool(need_g or need_b))
        dx = dw = None
        if need_x:
            dx = K.upconv3_dgrad(dconv, wd) if pre_up else K.conv3_igemm(dconv, wd)
        if need_w:
            if wgrad_side.active and ctx.wparam is not None and dconv.is_cuda:
                # on the weight-gradient stream; accumulated there and handed to .grad at the join (dw stays None here)
                wgrad_side.launch(ctx.wparam, (lambda: K.upconv3_wgrad(x, dconv)) if pre_up
                                  else (lambda: K.conv3_wgrad(x, dconv)), (x, dconv))
            else:
                dw = K.upconv3_wgrad(x, dconv) if pre_up else K.conv3_wgrad(x, dconv)
        return dx, dw, (dgamma if need_g else None), (dbeta if need_b else None), dres, None, None, None, None


def conv_bn_act(x, weight, gamma, beta, res, bn: BnState, slope: float, resample: int = K.RESAMPLE_NONE,
                pre_up: bool = False):
    return _ConvBnAct.apply(x, weight, gamma, beta, res, bn, slope, resample, pre_up)


class _Conv3(torch.autograd.Function):
    """Plain 3x3x3 convolution (no BatchNorm): y = conv3(x, weight).  Only used for the 1x1 projection shortcut of a
    stride-1 block that changes its channel count (models/models.py:28-43), which runs as the centre tap of a 3x3x3
    kernel -- no shipped block_setting takes that path, so it reuses the validated kernels instead of
    getting one of its own."""

    @staticmethod
    def forward(ctx, x, weight):
        wf, wd = _packed(weight, False)
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, wd)
        return K.conv3_igemm(x, wf)

    @staticmethod
    def backward(ctx, g):
        x, wd = ctx.saved_tensors
        g = g.contiguous()
        dx = K.conv3_igemm(g, wd) if ctx.needs_input_grad[0] else None
        dw = K.conv3_wgrad(x, g) if ctx.needs_input_grad[1] else None
        return dx, dw


def conv3(x, weight):
    return _Conv3.apply(x, weight)


class _StemBnAct(torch.autograd.Function):
    """x1 fp32 [N,D,H,W] -> NDHWC [N,D,H,W,C]; weight fp32 [C,T] (T = 27 or 1), bias [C]."""

    @staticmethod
    def forward(ctx, x1, weight, bias, gamma, beta, bn: BnState, slope: float, p: float):
        if bn.training:
            y, mean, invstd, scale, shift = K.c1_to_cn_bn(x1, weight, bias, gamma, beta, bn.running_mean, bn.running_var,
                                                          bn.num_batches_tracked, bn.momentum, bn.eps)
            _note_bn(bn, mean, invstd, y.numel() // y.shape[-1])
        else:
            y = K.c1_to_cn(x1, weight, bias)
            mean, invstd, scale, shift = _bn_coeffs(y, gamma, beta, bn)
        mask, seed, bits = None, 0, None
        p_eff = p if bn.training else 0.0
        if p_eff > 0.0:
            mask, seed = dropout_state.next()
            if mask is None and KEEP_BITS:
                # in-kernel Philox dropout: the forward kernel stores its keep decisions (1 bit per element) so the two
                # backward passes read 1/16 of a tensor instead of re-running Philox4x32-10
                bits = torch.empty(y.numel() // 8, dtype=torch.uint8, device=y.device)
        out = K.bn_act_fwd(y, scale, shift, None, slope, K.RESAMPLE_NONE, mask, p_eff, seed, keep_bits=bits)
        ctx.save_for_backward(x1, y, mean, invstd, gamma, beta, weight, mask, bits)
        ctx.cfg = (slope, p_eff, seed, bn.training)
        ctx.wparams = (_owning_param(weight), _owning_param(bias))
        return out

    @staticmethod
    def backward(ctx, g):
        x1, y, mean, invstd, gamma, beta, weight, mask, bits = ctx.saved_tensors
        slope, p, seed, training = ctx.cfg
        if not training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the reference hot path")
        need_x, need_w, need_bias, need_g, need_b = ctx.needs_input_grad[:5]
        dconv, _, dgamma, dbeta = K.bn_act_bwd(g.contiguous(), y, None, mean, invstd, gamma, beta, slope,
                                               K.RESAMPLE_NONE, mask, p, seed, need_dres=False,
                                               need_affine=bool(need_g or need_b), keep_bits=bits)
        dw = dbias = dx = None
        if need_w or need_bias:
            wp, bp = ctx.wparams
            if wgrad_side.active and dconv.is_cuda and wp is not None and bp is not None and need_w and need_bias:
                wgrad_side.launch((wp, bp), lambda: K.wgrad_c1(dconv, x1, weight.shape[1], flip=False)[:2], (dconv, x1))
            else:
                dw, dbias, _ = K.wgrad_c1(dconv, x1, weight.shape[1], flip=False)
        if need_x:
            dx = K.cn_to_c1(dconv, weight, None, flip=True, act=0)
        return (dx, dw if need_w else None, dbias if need_bias else None, dgamma if need_g else None,
                dbeta if need_b else None, None, None, None)


def stem_bn_act(x1, weight, bias, gamma, beta, bn: BnState, slope: float, p: float):
    return _StemBnAct.apply(x1, weight, bias, gamma, beta, bn, slope, p)


class _TailReluDrop(torch.autograd.Function):
    """a NDHWC [N,D,H,W,C] -> fp32 [N,D,H,W]; weight fp32 [C,27], bias [1]."""

    @staticmethod
    def forward(ctx, a, weight, bias, p: float, training: bool):
        mask, seed = (None, 0)
        p_eff = p if training else 0.0
        if p_eff > 0.0:
            mask, seed = dropout_state.next()
        out = K.cn_to_c1(a, weight, bias, flip=False, act=1, mask=mask, p=p_eff, seed=seed)
        ctx.save_for_backward(a, weight, out)
        ctx.p = p_eff
        ctx.wparams = (_owning_param(weight), _owning_param(bias))
        return out

    @staticmethod
    def backward(ctx, g):
        a, weight, out = ctx.saved_tensors
        need_a, need_w, need_b = ctx.needs_input_grad[:3]
        dy = K.relu_drop_bwd(g.contiguous(), out, ctx.p)
        da = K.c1_to_cn(dy, weight, None, flip=True) if need_a else None
        dw = db = None
        if need_w or need_b:
            wp, bp = ctx.wparams
            if wgrad_side.active and dy.is_cuda and wp is not None and bp is not None and need_w and need_b:
                def _tail_wgrad():
                    dw_, _, db_ = K.wgrad_c1(a, dy, weight.shape[1], flip=True)
                    return dw_, db_
                wgrad_side.launch((wp, bp), _tail_wgrad, (a, dy))
            else:
                dw, _, db = K.wgrad_c1(a, dy, weight.shape[1], flip=True)
        return da, (dw if need_w else None), (db if need_b else None), None, None


def tail_relu_drop(a, weight, bias, p: float, training: bool):
    return _TailReluDrop.apply(a, weight, bias, p, training)


class _Heads(torch.autograd.Function):
    """h NDHWC [N,d,h,w,C] -> (mu, logvar) fp32 [N,d,h,w]; weights fp32 [C,1], biases [1]."""

    @staticmethod
    def forward(ctx, h, w_mu, b_mu, w_var, b_var):
        mu = K.cn_to_c1(h, w_mu, b_mu)
        lv = K.cn_to_c1(h, w_var, b_var)
        ctx.save_for_backward(h, w_mu, w_var)
        return mu, lv

    @staticmethod
    def backward(ctx, dmu, dlv):
        h, w_mu, w_var = ctx.saved_tensors
        need_h, need_wm, need_bm, need_wv, need_bv = ctx.needs_input_grad
        dmu, dlv = dmu.contiguous(), dlv.contiguous()
        dh = None
        if need_h:
            dh = K.c1_to_cn(dmu, w_mu, None)
            dh = K.c1_to_cn(dlv, w_var, None, out=dh)
        dwm = dbm = dwv = dbv = None
        if need_wm or need_bm:
            dwm, _, dbm = K.wgrad_c1(h, dmu, 1)
        if need_wv or need_bv:
            dwv, _, dbv = K.wgrad_c1(h, dlv, 1)
        return (dh, dwm if need_wm else None, dbm if need_bm else None, dwv if need_wv else None,
                dbv if need_bv else None)


def heads(h, w_mu, b_mu, w_var, b_var):
    return _Heads.apply(h, w_mu, b_mu, w_var, b_var)


class _Reparam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar = mu.contiguous(), logvar.contiguous()
        if eps is None:
            # training path: eps ~ N(0,1) drawn inside the kernel (its Philox key = this call's seed + the device-side
            # epoch counter), returned for the backward pass
            z, eps = K.reparam_draw_fwd(mu, logvar, noise_state.next_seed())
        else:
            z = K.reparam_fwd(mu, logvar, eps)
        if isinstance(eps, torch.Tensor):
            ctx.save_for_backward(logvar, eps)
            ctx.eps_const = None
        else:
            ctx.save_for_backward(logvar)
            ctx.eps_const = float(eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        if ctx.eps_const is None:
            logvar, eps = ctx.saved_tensors
        else:
            (logvar,) = ctx.saved_tensors
            eps = ctx.eps_const
        dmu, dlv = K.reparam_bwd(dz.contiguous(), logvar, eps)
        return dmu, dlv, None


def reparameterize(mu, logvar, eps=None):
    """z = mu + eps*exp(0.5*logvar); eps is a tensor (injected, parity tests), a python float (validation, 0.1) or None:
    drawn ~ N(0,1) inside the kernel (the training path)."""
    return _Reparam.apply(mu, logvar, eps)


class _KlPerSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        ctx.shape = mu.shape
        mu2, lv2 = mu.reshape(mu.shape[0], -1).contiguous(), logvar.reshape(mu.shape[0], -1).contiguous()
        ctx.save_for_backward(mu2, lv2)
        return K.kl_persample_fwd(mu2, lv2)

    @staticmethod
    def backward(ctx, g):
        mu2, lv2 = ctx.saved_tensors
        dmu, dlv = K.kl_persample_bwd(mu2, lv2, g.contiguous())
        return dmu.reshape(ctx.shape), dlv.reshape(ctx.shape)


def kl_persample(mu, logvar):
    """[B] vector  -0.5*sum(1 + logvar - mu^2 - exp(logvar))."""
    return _KlPerSample.apply(mu, logvar)


class _MsePerSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.shape = x.shape
        x2, y2 = x.reshape(x.shape[0], -1).contiguous(), y.reshape(y.shape[0], -1).contiguous()
        ctx.save_for_backward(x2, y2)
        return K.mse_persample_fwd(x2, y2)

    @staticmethod
    def backward(ctx, g):
        x2, y2 = ctx.saved_tensors
        need_x, need_y = ctx.needs_input_grad
        dx, dy = K.mse_persample_bwd(x2, y2, g.contiguous(), bool(need_x), bool(need_y))
        return (dx.reshape(ctx.shape) if need_x else None), (dy.reshape(ctx.shape) if need_y else None)


def mse_persample(x, y):
    """[B] vector  sum_voxels (x - y)^2; gradients flow to both operands when both need them."""
    return _MsePerSample.apply(x, y)


class _IntroLossE(torch.autograd.Function):
    """lossE of utils/my_trainer.py:260-284 from the six per-sample vectors, one kernel forward / one backward.
    -> (lossE, mean r_real, mean k_real, exp_elbo_fake, exp_elbo_rec); only lossE is differentiable."""

    @staticmethod
    def forward(ctx, r_real, k_real, r_fake, k_fake, r_rec, k_rec, scale, b_rec, b_kl, b_neg):
        vs = [t.contiguous() for t in (r_real, k_real, r_fake, k_fake, r_rec, k_rec)]
        out = K.intro_loss_e_fwd(*vs, scale, b_rec, b_kl, b_neg)
        ctx.save_for_backward(vs[2], vs[3], vs[4], vs[5])
        ctx.hp = (scale, b_rec, b_kl, b_neg)
        outs = tuple(out[i] for i in range(5))
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g, *_unused):
        r_fake, k_fake, r_rec, k_rec = ctx.saved_tensors
        d = K.intro_loss_e_bwd(r_fake, k_fake, r_rec, k_rec, g.contiguous().reshape(1), *ctx.hp)
        need = ctx.needs_input_grad
        return tuple(d[i] if need[i] else None for i in range(6)) + (None, None, None, None)


def intro_loss_e(r_real, k_real, r_fake, k_fake, r_rec, k_rec, scale, b_rec, b_kl, b_neg):
    return _IntroLossE.apply(r_real, k_real, r_fake, k_fake, r_rec, k_rec, float(scale), float(b_rec), float(b_kl),
                             float(b_neg))


class _IntroLossD(torch.autograd.Function):
    """lossD of utils/my_trainer.py:301-321 from the five per-sample vectors.
    -> (lossD, mean r_real, mean k_rec, mean k_fake, mean r_rec_rec, mean r_fake_rec); only lossD is differentiable."""

    @staticmethod
    def forward(ctx, r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, scale, b_rec, b_kl, gamma_r):
        vs = [t.contiguous() for t in (r_real, k_rec, k_fake, r_rec_rec, r_fake_rec)]
        out = K.intro_loss_d_fwd(*vs, scale, b_rec, b_kl, gamma_r)
        ctx.batch = vs[0].numel()
        ctx.hp = (scale, b_rec, b_kl, gamma_r)
        outs = tuple(out[i] for i in range(6))
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g, *_unused):
        d = K.intro_loss_d_bwd(g.contiguous().reshape(1), ctx.batch, *ctx.hp)
        need = ctx.needs_input_grad
        return tuple(d[i] if need[i] else None for i in range(5)) + (None, None, None, None)


def intro_loss_d(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, scale, b_rec, b_kl, gamma_r):
    return _IntroLossD.apply(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, float(scale), float(b_rec), float(b_kl),
                             float(gamma_r))


class _Head1(torch.autograd.Function):
    """Single 1x1 head (ResNetEncoder.conv, models/models.py:105-108): NDHWC -> fp32 [N,d,h,w]."""

    @staticmethod
    def forward(ctx, h, w, b):
        ctx.save_for_backward(h, w)
        return K.cn_to_c1(h, w, b)

    @staticmethod
    def backward(ctx, dy):
        h, w = ctx.saved_tensors
        need_h, need_w, need_b = ctx.needs_input_grad
        dy = dy.contiguous()
        dh = K.c1_to_cn(dy, w, None) if need_h else None
        dw = db = None
        if need_w or need_b:
            dw, _, db = K.wgrad_c1(h, dy, 1)
        return dh, (dw if need_w else None), (db if need_b else None)


def head1(h, w, b):
    return _Head1.apply(h, w, b)

# ------------------------------------------------------------------------------------------------
# FC-latent variant (models/mymodel.py): Linear heads and the activation-carrying skip connection
# ------------------------------------------------------------------------------------------------
class _FcHead(torch.autograd.Function):
    """``x.view(B,-1) -> Linear`` (mymodel.py:140-141): h NDHWC [B,d,h,w,Cp] -> fp32 [B,J]; weight fp32 [J, c*d*h*w]."""

    @staticmethod
    def forward(ctx, h, weight, bias, c: int):
        xf = K.ndhwc_to_flat(h, c)
        w = weight.detach().contiguous()
        y = K.linear_fwd(xf, w, bias, relu=False)
        ctx.save_for_backward(xf, w)
        ctx.geom = (c, h.shape[-1], tuple(h.shape[1:4]))
        return y

    @staticmethod
    def backward(ctx, dy):
        xf, w = ctx.saved_tensors
        c, cp, grid = ctx.geom
        need_h, need_w, need_b = ctx.needs_input_grad[:3]
        dy = dy.contiguous()
        dh = K.flat_to_ndhwc(K.linear_dgrad(dy, w), c, cp, grid) if need_h else None
        dw = db = None
        if need_w or need_b:
            dw, db = K.linear_wgrad(xf, dy, need_bias=bool(need_b))
        return dh, (dw if need_w else None), db, None


def fc_head(h, weight, bias, c: int):
    return _FcHead.apply(h, weight, bias, c)


class _DfcHead(torch.autograd.Function):
    """``Linear -> ReLU -> view(B, C, d, h, w)`` (mymodel.py:150-153,:219): z fp32 [B,K] -> NDHWC [B,d,h,w,cp]."""

    @staticmethod
    def forward(ctx, z, weight, bias, c: int, cp: int, grid):
        z = z.reshape(z.shape[0], -1).contiguous().float()
        w = weight.detach().contiguous()
        y = K.linear_fwd(z, w, bias, relu=True)
        ctx.save_for_backward(z, y, w)
        ctx.c = c
        return K.flat_to_ndhwc(y, c, cp, grid)

    @staticmethod
    def backward(ctx, g):
        z, y, w = ctx.saved_tensors
        need_z, need_w, need_b = ctx.needs_input_grad[:3]
        dy = K.ndhwc_to_flat(g.contiguous(), ctx.c, gate=y)          # ReLU gradient folded into the layout change
        dz = K.linear_dgrad(dy, w) if need_z else None
        dw = db = None
        if need_w or need_b:
            dw, db = K.linear_wgrad(z, dy, need_bias=bool(need_b))
        return dz, (dw if need_w else None), db, None, None, None


def dfc_head(z, weight, bias, c: int, cp: int, grid):
    return _DfcHead.apply(z, weight, bias, c, cp, grid)


class _AddAct(torch.autograd.Function):
    """LeakyReLU(a + b) where b already went through its own activation (mymodel.py:135-136)."""

    @staticmethod
    def forward(ctx, a, b, slope: float):
        out = K.add_act_fwd(a.contiguous(), b.contiguous(), slope)
        ctx.save_for_backward(out)
        ctx.slope = slope
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        dz = K.add_act_bwd(g.contiguous(), out, ctx.slope)
        return (dz if ctx.needs_input_grad[0] else None), (dz if ctx.needs_input_grad[1] else None), None


def add_act(a, b, slope: float):
    return _AddAct.apply(a, b, slope)



# ----------------------------------------------------------------------------------------------
# reparameterisation noise: drawn inside the kernel by default (functional.manual_seed keys it), or an
# injected sequence of eps tensors for parity tests
# ----------------------------------------------------------------------------------------------
class _NoiseState:
    """``eps_feed``: optional iterator of injected eps tensors (parity tests).  Otherwise every reparameterisation call of
    a step gets its own Philox key (base seed, call index; the kernels mix in the device-side epoch counter that
    ``begin_step`` advances, exactly as for dropout), so a captured CUDA graph draws fresh noise on every replay."""

    def __init__(self):
        self.eps_feed = None
        self.base_seed = 0x0E9500E95
        self.call = 0

    def next_seed(self) -> int:
        self.call += 1
        return ((self.base_seed * 0xD6E8FEB86659FD93 + self.call * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF) | 1


noise_state = _NoiseState()


def draw_eps(like: torch.Tensor):
    """-> the injected eps tensor of this call, or None (= the kernel draws it; on CPU tensors: torch.randn_like)."""
    if noise_state.eps_feed is not None:
        return next(noise_state.eps_feed).to(like.device, like.dtype).reshape(like.shape)
    return None
