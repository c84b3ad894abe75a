"""Drop-in ``nn.Module`` shells for the reference's ``models/mymodel.py`` -- the FC-latent Soft-IntroVAE variant
(vector latent ``[B, z_ch]``; SURVEY.md section 8f NEXT-1, BASELINE config 2 ``600z_main.py``) -- whose
forward/backward run on libsivae.so.

Same class names, constructor signatures, attribute tree, parameter-creation order and ``state_dict`` keys as the
reference (models/mymodel.py:51-290): ``encoder.block1 .. block8 / fc``, ``decoder.dfc / block1 / block2u / block3 /
block4u / block5u / block6u / last_block``; the sub-modules are stock ``nn.Conv3d`` / ``nn.BatchNorm3d`` /
``nn.Linear`` parameter holders and only ``forward`` is replaced.

Differences from ``models.py`` that matter to the kernels:
  * every convolution has ``bias=True`` and (except ``last_block``) is followed by a BatchNorm3d.  In train mode the
    bias cancels in the normalisation; the fused units therefore run the bias-free convolution and the bias only
    enters the running mean (``running_mean += momentum * bias``, one multi-tensor update per pass).  In eval mode it
    is folded into the running mean handed to the kernels.  The bias gradient is mathematically zero; the reference
    produces round-off noise there (tests/test_oracle_vs_golden.py::test_fc_soft_intro_step bounds it at 1e-3 of the
    weight gradient), this build leaves ``bias.grad = None`` so Adam skips those 1-D tensors.
  * the encoder's first skip (``Leakyrelu5(x + block5(x))``, :135-136) adds a branch that already carries its own
    LeakyReLU, so it is an element-wise add + activation kernel rather than the fused BN-residual unit; the other
    skips (block7, decoder block1 / block3) are ``LeakyReLU(x + BN(conv))`` and use the fused unit.
  * ``block8`` is constructed but never executed (:108-117): its parameters keep ``grad is None``.
  * the Linear heads work on the NCDHW-flattened 5x6x5 map (:125,:140,:151,:219): weight-streaming kernels in
    csrc/linear.cu, fp32 weights exactly as stored.

The reference hard-codes the 5x6x5 grid (80x96x80 inputs).  ``latent_grid`` is the one added constructor keyword
(default ``(5, 6, 5)``): parity tests use smaller grids; inputs must be ``16 * latent_grid``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn

from . import functional as F
from . import kernels as K
from .models import BaseVAE, _BnBinding, _conv3_weight, _cpad, _pad_dim

SLOPE = 0.2


def _cbl(cin, cout, act=True):
    """Conv3d(bias=True) -> BatchNorm3d -> [LeakyReLU(0.2)] exactly as spelled out in mymodel.py:55-58."""
    mods = [nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm3d(cout)]
    if act:
        mods.append(nn.LeakyReLU(0.2, inplace=True))
    return mods


class _BiasIntoRunningMean:
    """Collects (running_mean, bias) pairs of one pass and applies ``running_mean += momentum * bias`` in a single
    multi-tensor launch (train mode only)."""

    def __init__(self):
        self.rms, self.biases, self.mom = [], [], None

    def add(self, bn: nn.BatchNorm3d, bias: torch.Tensor):
        mom = 0.1 if bn.momentum is None else bn.momentum
        if self.mom is not None and mom != self.mom:
            self.flush()
        self.mom = mom
        self.rms.append(bn.running_mean)
        self.biases.append(bias.detach())

    def flush(self):
        if self.rms:
            with torch.no_grad():
                torch._foreach_add_(self.rms, self.biases, alpha=self.mom)
        self.rms, self.biases = [], []


def _conv_bn_act(x, conv: nn.Conv3d, bn: nn.BatchNorm3d, pend: _BiasIntoRunningMean, res=None,
                 resample: int = K.RESAMPLE_NONE, pre_up: bool = False, slope: float = SLOPE):
    b = _BnBinding(bn)
    gamma, beta = b.params()
    state = b.state()
    if conv.bias is not None and not bn.training:
        state.running_mean = state.running_mean - _pad_dim(conv.bias.detach(), 0, b.cp)
    out = F.conv_bn_act(x, _conv3_weight(conv), gamma, beta, res, state, slope, resample, pre_up)
    b.commit()
    if conv.bias is not None and bn.training:
        if state.defer is not None:
            F.note_bias_into_running_mean(conv.bias)     # side-stream pass: rides on the deferred running-mean update
        else:
            pend.add(bn, conv.bias)
    return out


class ResNetVAEencoder(nn.Module):
    """Reference: models/mymodel.py:51-143; forward returns (mu, logvar), each [B, z_ch]."""

    def __init__(self, first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid: Tuple[int, int, int] = (5, 6, 5)):
        super().__init__()
        self.forth_ch = forth_ch
        self.latent_grid = tuple(latent_grid)
        self.block1 = nn.Sequential(*_cbl(1, first_ch), *_cbl(first_ch, first_ch))
        self.block2 = nn.Sequential(*_cbl(first_ch, first_ch), *_cbl(first_ch, second_ch))
        self.block3 = nn.Sequential(*_cbl(second_ch, second_ch), *_cbl(second_ch, third_ch))
        self.block4short = nn.Sequential(*_cbl(third_ch, third_ch))
        self.block5 = nn.Sequential(*_cbl(third_ch, third_ch))
        self.block6 = nn.Sequential(*_cbl(third_ch, third_ch), nn.AvgPool3d(kernel_size=2), *_cbl(third_ch, forth_ch))
        self.block7 = nn.Sequential(*_cbl(forth_ch, forth_ch), *_cbl(forth_ch, forth_ch, act=False))
        self.block8 = nn.Sequential(*_cbl(third_ch, third_ch), *_cbl(third_ch, forth_ch))   # never run (:108-117)
        self.pool1 = nn.AvgPool3d(kernel_size=2)
        self.pool2 = nn.AvgPool3d(kernel_size=2)
        self.pool3 = nn.AvgPool3d(kernel_size=2)
        self.pool4 = nn.AvgPool3d(kernel_size=2)
        self.Leakyrelu5 = nn.LeakyReLU(0.2, inplace=True)
        self.Leakyrelu7 = nn.LeakyReLU(0.2, inplace=True)
        g = self.latent_grid
        self.fc = nn.Linear(forth_ch * g[0] * g[1] * g[2], z_ch * 2)

    def forward(self, x: torch.Tensor):
        g = self.latent_grid
        if x.dim() != 5 or x.shape[1] != 1 or tuple(x.shape[2:]) != tuple(16 * v for v in g):
            raise ValueError(f"expected [B,1,{16 * g[0]},{16 * g[1]},{16 * g[2]}], got {tuple(x.shape)}")
        pend = _BiasIntoRunningMean()
        pool = K.RESAMPLE_AVGPOOL2
        x1 = x.reshape(x.shape[0], x.shape[2], x.shape[3], x.shape[4]).contiguous().float()
        conv, bn = self.block1[0], self.block1[1]
        c, cp = conv.out_channels, _cpad(conv.out_channels)
        b = _BnBinding(bn)
        gamma, beta = b.params()
        # the stem kernel takes the bias itself (it is part of the hi/lo-split GEMM), so nothing is pending for it
        h = F.stem_bn_act(x1, _pad_dim(conv.weight.reshape(c, 27), 0, cp), _pad_dim(conv.bias, 0, cp), gamma, beta,
                          b.state(), SLOPE, 0.0)
        b.commit()
        h = _conv_bn_act(h, self.block1[3], self.block1[4], pend, resample=pool)          # + pool1 (:129)
        h = _conv_bn_act(h, self.block2[0], self.block2[1], pend)
        h = _conv_bn_act(h, self.block2[3], self.block2[4], pend, resample=pool)          # + pool2 (:131)
        h = _conv_bn_act(h, self.block3[0], self.block3[1], pend)
        h = _conv_bn_act(h, self.block3[3], self.block3[4], pend, resample=pool)          # + pool3 (:133)
        h = _conv_bn_act(h, self.block4short[0], self.block4short[1], pend)
        r = _conv_bn_act(h, self.block5[0], self.block5[1], pend)
        h = F.add_act(h, r, SLOPE)                                                         # :136
        h = _conv_bn_act(h, self.block6[0], self.block6[1], pend, resample=pool)          # block6[3]
        h = _conv_bn_act(h, self.block6[4], self.block6[5], pend)
        r = _conv_bn_act(h, self.block7[0], self.block7[1], pend)
        h = _conv_bn_act(r, self.block7[3], self.block7[4], pend, res=h)                  # :138-139
        pend.flush()
        y = F.fc_head(h, self.fc.weight, self.fc.bias, self.forth_ch)
        mu, logvar = y.chunk(2, dim=1)
        return mu, logvar


class ResNetDecoder(nn.Module):
    """Reference: models/mymodel.py:146-230; [B, z_ch] -> [B,1,80,96,80]."""

    def __init__(self, first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid: Tuple[int, int, int] = (5, 6, 5)):
        super().__init__()
        self.forth_ch = forth_ch
        self.latent_grid = tuple(latent_grid)
        g = self.latent_grid
        self.dfc = nn.Sequential(nn.Linear(z_ch, forth_ch * g[0] * g[1] * g[2]), nn.ReLU(True))
        up = lambda: nn.Upsample(scale_factor=2, mode="nearest")  # noqa: E731
        self.block1 = nn.Sequential(*_cbl(forth_ch, forth_ch), *_cbl(forth_ch, forth_ch, act=False))
        self.block2u = nn.Sequential(*_cbl(forth_ch, forth_ch), up(), *_cbl(forth_ch, third_ch))
        self.block3 = nn.Sequential(*_cbl(third_ch, third_ch), *_cbl(third_ch, third_ch, act=False))
        self.block4u = nn.Sequential(*_cbl(third_ch, third_ch), up(), *_cbl(third_ch, second_ch))
        self.block5u = nn.Sequential(*_cbl(second_ch, second_ch), up(), *_cbl(second_ch, first_ch))
        self.block6u = nn.Sequential(*_cbl(first_ch, first_ch), up(), *_cbl(first_ch, first_ch))
        self.last_block = nn.Sequential(nn.Conv3d(first_ch, 1, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU())
        self.dLeakyrelu1 = nn.LeakyReLU(0.2, inplace=True)
        self.dLeakyrelu2 = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        pend = _BiasIntoRunningMean()
        lin = self.dfc[0]
        y = F.dfc_head(z, lin.weight, lin.bias, self.forth_ch, _cpad(self.forth_ch), self.latent_grid)
        r = _conv_bn_act(y, self.block1[0], self.block1[1], pend)
        y = _conv_bn_act(r, self.block1[3], self.block1[4], pend, res=y)                  # :220-221
        y = _conv_bn_act(y, self.block2u[0], self.block2u[1], pend)
        y = _conv_bn_act(y, self.block2u[4], self.block2u[5], pend, pre_up=True)
        r = _conv_bn_act(y, self.block3[0], self.block3[1], pend)
        y = _conv_bn_act(r, self.block3[3], self.block3[4], pend, res=y)                  # :223-224
        for blk in (self.block4u, self.block5u, self.block6u):
            y = _conv_bn_act(y, blk[0], blk[1], pend)
            y = _conv_bn_act(y, blk[4], blk[5], pend, pre_up=True)
        pend.flush()
        tconv = self.last_block[0]
        ci = tconv.in_channels
        out = F.tail_relu_drop(y, _pad_dim(tconv.weight.reshape(ci, 27), 0, _cpad(ci)), tconv.bias, 0.0, False)
        return out.unsqueeze(1)


class ResNetVAE(BaseVAE):
    """Reference: models/mymodel.py:233-248; forward returns (x_re, mu, logvar)."""

    def __init__(self, first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid=(5, 6, 5)) -> None:
        super().__init__()
        self.encoder = ResNetVAEencoder(first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid)
        self.decoder = ResNetDecoder(first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid)

    def reparamenterize(self, mu, logvar):  # sic: the reference's spelling (mymodel.py:239)
        return F.reparameterize(mu, logvar, F.draw_eps(mu))

    def forward(self, x):
        mu, logvar = self.encoder(x)
        z = self.reparamenterize(mu, logvar)
        x_re = self.decoder(z)
        return x_re, mu, logvar


class SoftIntroVAE(nn.Module):
    """Reference: models/mymodel.py:256-290; forward returns (mu, logvar, z, x_re)."""

    def __init__(self, first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid=(5, 6, 5)) -> None:
        super().__init__()
        self.encoder = ResNetVAEencoder(first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid)
        self.decoder = ResNetDecoder(first_ch, second_ch, third_ch, forth_ch, z_ch, latent_grid)
        self.z_ch = z_ch

    def two_stream_ok(self) -> bool:
        """Independent passes of a training iteration may run on two CUDA streams (trainer._fork_join)."""
        return True

    def prepack(self):
        """bf16 packs of every unpadded 3x3x3 weight on the current stream before two passes fork (padded weights are
        re-packed per call from their temporaries, each pass its own)."""
        e, d = self.encoder, self.decoder
        plain = [e.block1[3], e.block2[0], e.block2[3], e.block3[0], e.block3[3], e.block4short[0], e.block5[0],
                 e.block6[0], e.block6[4], e.block7[0], e.block7[3], d.block1[0], d.block1[3], d.block2u[0], d.block3[0],
                 d.block3[3], d.block4u[0], d.block5u[0], d.block6u[0]]
        up = [d.block2u[4], d.block4u[4], d.block5u[4], d.block6u[4]]
        for conv, pre_up in [(c, False) for c in plain] + [(c, True) for c in up]:
            w = _conv3_weight(conv)
            if w is conv.weight:
                F._packed(w, pre_up)

    def reparameterize(self, mu, logvar):
        return F.reparameterize(mu, logvar, F.draw_eps(mu))

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
        z = z.view(32, 1, 5, 6, 5)  # hard-coded in the reference (mymodel.py:284); the decoder flattens it again
        return self.decode(z, y_cond=y_cond)

    def sample_with_noise(self, num_samples=1, device=torch.device("cpu"), y_cond=None):
        # the reference reads an undefined ``self.z_dim`` here (mymodel.py:289); kept
        z = torch.randn(num_samples, self.z_dim).to(device)
        return self.decode(z, y_cond=y_cond)


__all__: List[str] = ["ResNetVAEencoder", "ResNetDecoder", "ResNetVAE", "SoftIntroVAE"]
