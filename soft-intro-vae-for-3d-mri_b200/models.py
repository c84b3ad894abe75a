"""Drop-in ``nn.Module`` shells for the reference's ``models/models.py`` (LeakyReLU(0.2) + Dropout
variant) whose forward/backward run on libsivae.so.

Same class names, constructor signatures, attribute tree, parameter-creation order (so that
``torch.manual_seed(s); SoftIntroVAE(...)`` initialises identically) and ``state_dict`` keys as the
reference (models/models.py:8-300): the sub-modules are stock ``nn.Conv3d`` / ``nn.BatchNorm3d``
*parameter holders* (exact types, so ``init_weights_he`` -- utils/my_trainer.py:511-514 -- and
checkpoints keep working); only ``forward`` is replaced at block / encoder / decoder level by calls
into the fused units of ``functional.py``.

Inputs / outputs stay NCDHW fp32 with C = 1 ([B,1,D,H,W], latent [B,1,D/8,H/8,W/8] for three
stride-2 stages); activations in between are NDHWC bf16 and never leave the device.  Channel
counts that are not a multiple of 64 are zero-padded to the next multiple (the tcgen05 tiles are
64 channels wide); the headline network (64/128/256) needs no padding.

Reference quirks kept on purpose: the residual is applied only when
stride == 1 and the 1x1 projection ``shortcut`` conv is created but never executed (Q1);
``encoder.conv`` exists but is unused by the VAE encoder (Q2); the decoder output passes
ReLU then Dropout(0.35) in train mode (Q3); ``SoftIntroVAE.sample`` keeps its hard-coded view (Q19).
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F
from . import kernels as K

CH_ALIGN = 64


def _cpad(c: int) -> int:
    return (c + CH_ALIGN - 1) // CH_ALIGN * CH_ALIGN


def _pad_dim(t: torch.Tensor, dim: int, size: int, value: float = 0.0) -> torch.Tensor:
    """Zero-pad dimension ``dim`` of ``t`` up to ``size`` (differentiable; no-op when already there)."""
    cur = t.shape[dim]
    if cur == size:
        return t
    pad = [0, 0] * (t.dim() - 1 - dim) + [0, size - cur]
    return TF.pad(t, pad, value=value)


class _BnBinding:
    """Binds an nn.BatchNorm3d holder to the (possibly channel-padded) fused kernels."""

    def __init__(self, bn: nn.BatchNorm3d):
        self.bn = bn
        self.c = bn.num_features
        self.cp = _cpad(self.c)
        self._tmp = None

    def params(self):
        return _pad_dim(self.bn.weight, 0, self.cp), _pad_dim(self.bn.bias, 0, self.cp)

    def state(self) -> F.BnState:
        bn = self.bn
        mom = 0.1 if bn.momentum is None else bn.momentum
        if F.bn_defer.records is not None and bn.training:
            # side-stream pass (trainer._fork_join): the running statistics are updated after the streams have joined
            return F.BnState(None, None, None, True, mom, bn.eps,
                             defer=(bn.running_mean, bn.running_var, bn.num_batches_tracked))
        if self.c == self.cp:
            return F.BnState(bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.training, mom, bn.eps)
        rm = _pad_dim(bn.running_mean, 0, self.cp)
        rv = _pad_dim(bn.running_var, 0, self.cp, 1.0)
        self._tmp = (rm, rv)
        return F.BnState(rm, rv, bn.num_batches_tracked, bn.training, mom, bn.eps)

    def commit(self):
        """Copy padded running statistics back into the holder (padded path only)."""
        if self._tmp is not None and self.bn.training:
            with torch.no_grad():
                self.bn.running_mean.copy_(self._tmp[0][: self.c])
                self.bn.running_var.copy_(self._tmp[1][: self.c])
        self._tmp = None


def _conv3_weight(conv: nn.Conv3d) -> torch.Tensor:
    """[Co,Ci,3,3,3] zero-padded to 64-multiples on both channel dims."""
    w = conv.weight
    return _pad_dim(_pad_dim(w, 0, _cpad(w.shape[0])), 1, _cpad(w.shape[1]))


def _fused_conv_bn_act(x, conv: nn.Conv3d, bn: nn.BatchNorm3d, res, slope: float, resample: int, pre_up: bool = False):
    b = _BnBinding(bn)
    gamma, beta = b.params()
    out = F.conv_bn_act(x, _conv3_weight(conv), gamma, beta, res, b.state(), slope, resample, pre_up)
    b.commit()
    return out


def _act_module(slope: float) -> nn.Module:
    return nn.LeakyReLU(slope, inplace=True) if slope != 0.0 else nn.ReLU(inplace=True)


class BuildingBlock(nn.Module):
    """conv3 -> BN -> act -> AvgPool(stride) -> conv3 -> BN, (+x iff stride == 1), act.
    Reference: models/models.py:8-43."""

    _upsample = False

    def __init__(self, in_ch, out_ch, stride, bias=False, slope: float = 0.2, with_dropout_attr: bool = True):
        super().__init__()
        self.res = stride == 1
        self.stride = stride
        self.slope = slope
        if stride not in (1, 2):
            raise ValueError("libsivae fuses AvgPool3d/Upsample with factor 1 or 2 only")
        if bias:
            raise ValueError("block convolutions are bias-free in every shipped configuration (models.py:9)")
        # creation order mirrors the reference so seeded construction matches parameter for parameter
        self.shortcut = self._shortcut(in_ch, out_ch)
        if with_dropout_attr:
            self.dropout = nn.Dropout(p=0.25)  # never called
        self.relu = _act_module(slope)
        mid = in_ch if self._upsample else out_ch
        resample = nn.Upsample(scale_factor=stride) if self._upsample else nn.AvgPool3d(kernel_size=stride)
        self.block = nn.Sequential(
            nn.Conv3d(in_ch, mid, kernel_size=3, stride=1, padding=1, bias=bias),
            nn.BatchNorm3d(mid),
            _act_module(slope),
            resample,
            nn.Conv3d(mid, out_ch, kernel_size=3, stride=1, padding=1, bias=bias),
            nn.BatchNorm3d(out_ch),
        )

    def _shortcut(self, in_ch, out_ch):
        if in_ch != out_ch:
            return self._projection(in_ch, out_ch)
        return lambda x: x

    def _projection(self, channel_in, channel_out):
        return nn.Conv3d(channel_in, channel_out, kernel_size=1, stride=1, padding=0)

    def prepack(self):
        """Make sure the bf16 weight packs of both convolutions exist (on the CURRENT stream) -- called before two passes
        through this block are issued on two streams, so neither of them launches a pack kernel the other one needs."""
        for conv, pre_up in ((self.block[0], False), (self.block[4], self.stride == 2 and self._upsample)):
            w = _conv3_weight(conv)
            if w is conv.weight:          # channel-padded weights are re-packed per call from their temporaries
                F._packed(w, pre_up)

    def forward(self, x):
        """x, result: NDHWC activations (bf16 on the CUDA path)."""
        if self.res and not isinstance(self.shortcut, nn.Module):
            res = x
        elif self.res:
            # stride-1 block that changes its channel count: the reference runs the 1x1 projection conv here
            # (models/models.py:32-35,39-41).  No shipped block_setting does, so the projection reuses the
            # 3x3x3 kernels with its weight as the centre tap (27x its FLOPs) and adds the bias in torch.
            sc = self.shortcut
            w1 = _pad_dim(_pad_dim(sc.weight, 0, _cpad(sc.out_channels)), 1, _cpad(sc.in_channels))
            w3 = TF.pad(w1, (1, 1, 1, 1, 1, 1))
            res = F.conv3(x, w3)
            if sc.bias is not None:
                res = res + _pad_dim(sc.bias, 0, _cpad(sc.out_channels)).to(res.dtype)
        else:
            res = None
        # AvgPool3d(2) is fused behind the first convolution's BN/activation; Upsample(2) is folded into the
        # second convolution itself (never materialised)
        mode = K.RESAMPLE_AVGPOOL2 if (self.stride == 2 and not self._upsample) else K.RESAMPLE_NONE
        a = _fused_conv_bn_act(x, self.block[0], self.block[1], None, self.slope, mode)
        return _fused_conv_bn_act(a, self.block[4], self.block[5], res, self.slope, K.RESAMPLE_NONE,
                                  pre_up=(self.stride == 2 and self._upsample))


class UpsampleBuildingkBlock(BuildingBlock):
    """conv3(in->in) -> BN -> act -> Upsample(stride, nearest) -> conv3(in->out) -> BN, (+x), act.
    Reference: models/models.py:46-80 (the class name keeps the reference's spelling)."""

    _upsample = True


class ResNetEncoder(nn.Module):
    """Reference: models/models.py:83-108."""

    _slope = 0.2
    _p_stem = 0.35
    _block_dropout_attr = True

    def __init__(self, in_ch, block_setting):
        super().__init__()
        self.block_setting = block_setting
        if self._block_dropout_attr:
            self.dropout = nn.Dropout(p=0.25)  # never called
        self.in_ch = in_ch
        last = 1
        stem = [nn.Conv3d(1, in_ch, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm3d(in_ch),
                _act_module(self._slope)]
        if self._p_stem > 0:
            stem.append(nn.Dropout(p=self._p_stem))
        blocks = [nn.Sequential(*stem)]
        for line in self.block_setting:
            c, n, s = line[0], line[1], line[2]
            for i in range(n):
                stride = s if i == 0 else 1
                blocks.append(nn.Sequential(BuildingBlock(in_ch, c, stride, slope=self._slope,
                                                          with_dropout_attr=self._block_dropout_attr)))
                in_ch = c
        self.inner_ch = in_ch
        self.blocks = nn.Sequential(*blocks)
        self.conv = nn.Sequential(nn.Conv3d(in_ch, last, kernel_size=1, stride=1, bias=True))

    # -- fused path -------------------------------------------------------------------------
    def features(self, x: torch.Tensor) -> torch.Tensor:
        """[B,1,D,H,W] fp32 -> NDHWC activations of the last block."""
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected [B,1,D,H,W], got {tuple(x.shape)}")
        x1 = x.reshape(x.shape[0], x.shape[2], x.shape[3], x.shape[4]).contiguous().float()
        stem = self.blocks[0]
        conv, bn = stem[0], stem[1]
        c, cp = conv.out_channels, _cpad(conv.out_channels)
        b = _BnBinding(bn)
        gamma, beta = b.params()
        p = stem[3].p if (len(stem) > 3 and stem[3].training) else 0.0
        h = F.stem_bn_act(x1, _pad_dim(conv.weight.reshape(c, 27), 0, cp), _pad_dim(conv.bias, 0, cp), gamma, beta,
                          b.state(), self._slope, p)
        b.commit()
        for blk in list(self.blocks)[1:]:
            h = blk[0](h)
        return h

    def _head(self, h, conv: nn.Conv3d):
        c = conv.in_channels
        return F.head1(h, _pad_dim(conv.weight.reshape(c, 1), 0, _cpad(c)), conv.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._head(self.features(x), self.conv[0]).unsqueeze(1)


class ResNetDecoder(nn.Module):
    """Reference: models/models.py:110-145."""

    _slope = 0.2
    _p_stem = 0.25
    _p_tail = 0.35
    _block_dropout_attr = True

    def __init__(self, encoder: ResNetEncoder, blocks=None):
        super().__init__()
        if self._block_dropout_attr:
            self.dropout = nn.Dropout(p=0.25)  # never called
        if blocks is not None:
            raise NotImplementedError("custom decoder stems are not used by any entry script")
        last = encoder.block_setting[-1][0]
        stem = [nn.Conv3d(1, last, 1, 1, bias=True), nn.BatchNorm3d(last), _act_module(self._slope)]
        if self._p_stem > 0:
            stem.append(nn.Dropout(p=self._p_stem))
        blocks = [nn.Sequential(*stem)]
        in_ch = last
        for i in range(len(encoder.block_setting)):
            if i == len(encoder.block_setting) - 1:
                nc = encoder.in_ch
            else:
                nc = encoder.block_setting[::-1][i + 1][0]
            c, n, s = encoder.block_setting[::-1][i]
            for j in range(n):
                stride = s if j == n - 1 else 1
                c = nc if j == n - 1 else c
                blocks.append(nn.Sequential(UpsampleBuildingkBlock(in_ch, c, stride, slope=self._slope,
                                                                   with_dropout_attr=self._block_dropout_attr)))
                in_ch = c
        tail = [nn.Conv3d(in_ch, 1, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU()]
        if self._p_tail > 0:
            tail.append(nn.Dropout(p=self._p_tail))
        blocks.append(nn.Sequential(*tail))
        self.blocks = nn.Sequential(*blocks)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """[B,1,d,h,w] fp32 -> [B,1,8d,8h,8w] fp32 (for three stride-2 stages)."""
        if z.dim() != 5 or z.shape[1] != 1:
            raise ValueError(f"expected [B,1,d,h,w], got {tuple(z.shape)}")
        z1 = z.reshape(z.shape[0], z.shape[2], z.shape[3], z.shape[4]).contiguous().float()
        mods = list(self.blocks)
        stem, tail = mods[0], mods[-1]
        conv, bn = stem[0], stem[1]
        c, cp = conv.out_channels, _cpad(conv.out_channels)
        b = _BnBinding(bn)
        gamma, beta = b.params()
        p = stem[3].p if (len(stem) > 3 and stem[3].training) else 0.0
        h = F.stem_bn_act(z1, _pad_dim(conv.weight.reshape(c, 1), 0, cp), _pad_dim(conv.bias, 0, cp), gamma, beta,
                          b.state(), self._slope, p)
        b.commit()
        for blk in mods[1:-1]:
            h = blk[0](h)
        tconv = tail[0]
        ci = tconv.in_channels
        p_tail = tail[2].p if len(tail) > 2 else 0.0
        training = tail[2].training if len(tail) > 2 else False
        out = F.tail_relu_drop(h, _pad_dim(tconv.weight.reshape(ci, 27), 0, _cpad(ci)), tconv.bias, p_tail, training)
        return out.unsqueeze(1)


class BaseEncoder(nn.Module):
    def __init__(self) -> None:
        super().__init__()


class BaseDecoder(nn.Module):
    def __init__(self) -> None:
        super().__init__()


class BaseCAE(nn.Module):
    """Reference: models/models.py:155-168."""

    def __init__(self) -> None:
        super().__init__()
        self.encoder = BaseEncoder()
        self.decoder = BaseDecoder()

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)

    def forward(self, x):
        z = self.encode(x)
        return self.decode(z), z


class ResNetCAE(BaseCAE):
    """Reference: models/models.py:171-187."""

    _enc, _dec = ResNetEncoder, ResNetDecoder

    def __init__(self, in_ch, block_setting) -> None:
        super().__init__()
        self.encoder = self._enc(in_ch=in_ch, block_setting=block_setting)
        self.decoder = self._dec(self.encoder)

    def forward(self, x):
        return self.decoder(self.encoder(x))


class BaseVAE(nn.Module):
    """Reference: models/models.py:190-210."""

    def __init__(self) -> None:
        super().__init__()
        self.encoder = BaseEncoder()
        self.decoder = BaseDecoder()

    def encode(self, x):
        mu, logvar = self.encoder(x)
        return mu, logvar

    def decode(self, vec):
        return self.decoder(vec)

    def reparameterize(self, mu, logvar) -> torch.Tensor:
        return F.reparameterize(mu, logvar, F.draw_eps(mu))

    def forward(self, x):
        mu, logvar = self.encode(x)
        vec = self.reparameterize(mu, logvar)
        x_hat = self.decode(vec)
        return x_hat, vec, mu, logvar


class VAEResNetEncoder(ResNetEncoder):
    """Reference: models/models.py:213-223; forward returns (mu, logvar)."""

    def __init__(self, in_ch, block_setting) -> None:
        super().__init__(in_ch, block_setting)
        self.mu = nn.Conv3d(self.inner_ch, 1, kernel_size=1, stride=1, bias=True)
        self.var = nn.Conv3d(self.inner_ch, 1, kernel_size=1, stride=1, bias=True)

    def forward(self, x: torch.Tensor):
        h = self.features(x)
        c, cp = self.inner_ch, _cpad(self.inner_ch)
        mu, lv = F.heads(h, _pad_dim(self.mu.weight.reshape(c, 1), 0, cp), self.mu.bias,
                         _pad_dim(self.var.weight.reshape(c, 1), 0, cp), self.var.bias)
        return mu.unsqueeze(1), lv.unsqueeze(1)


class ResNetVAE(BaseVAE):
    """Reference: models/models.py:226-241; forward returns (x_re, mu, logvar)."""

    _enc, _dec = VAEResNetEncoder, ResNetDecoder

    def __init__(self, in_ch, block_setting) -> None:
        super().__init__()
        self.encoder = self._enc(in_ch=in_ch, block_setting=block_setting)
        self.decoder = self._dec(self.encoder)

    def reparamenterize(self, mu, logvar):  # sic: the reference's spelling (models.py:232)
        return F.reparameterize(mu, logvar, F.draw_eps(mu))

    def forward(self, x):
        mu, logvar = self.encoder(x)
        z = self.reparamenterize(mu, logvar)
        x_re = self.decoder(z)
        return x_re, mu, logvar


class SoftIntroVAE(nn.Module):
    """Reference: models/models.py:257-300; forward returns (mu, logvar, z, x_re)."""

    def __init__(self, in_ch, block_setting) -> None:
        super().__init__()
        self.encoder = VAEResNetEncoder(in_ch=in_ch, block_setting=block_setting)
        self.decoder = ResNetDecoder(self.encoder)

    def two_stream_ok(self) -> bool:
        """Independent passes of a training iteration may run on two CUDA streams (trainer._fork_join)."""
        return True

    def prepack(self):
        for m in self.modules():
            if isinstance(m, BuildingBlock):
                m.prepack()

    def reparameterize(self, mu, logvar, val_flag=False):
        # train: eps ~ N(0,1) drawn inside the reparameterisation kernel (or injected); validation: the constant 0.1 (Q9)
        eps = F.draw_eps(mu) if val_flag is False else 0.1
        return F.reparameterize(mu, logvar, eps)

    def forward(self, x):
        mu, logvar = self.encoder(x)
        z = self.reparameterize(mu, logvar)
        x_re = self.decoder(z)
        return mu, logvar, z, x_re

    def encode(self, x, o_cond=None):
        mu, logvar = self.encoder(x)
        return mu, logvar

    def decode(self, z, y_cond=None):
        return self.decoder(z)

    def sample(self, z, y_cond=None):
        z = z.view(32, 1, 5, 6, 5)  # hard-coded in the reference (models.py:294)
        return self.decode(z, y_cond=y_cond)

    def sample_with_noise(self, num_samples=1, device=torch.device("cpu"), y_cond=None):
        # the reference reads an undefined ``self.z_dim`` here (models.py:299)
        z = torch.randn(num_samples, self.z_dim).to(device)
        return self.decode(z, y_cond=y_cond)
