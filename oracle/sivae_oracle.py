"""Torch-fp32 restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

A *functional* re-statement (plain ``torch.nn.functional`` calls over a
``state_dict``) of the reference's Soft-IntroVAE / ResNetVAE forward passes, loss
functions and one Soft-IntroVAE training iteration.  It is device-agnostic: on the
authoring box it runs on CPU, on the GPU box the parity tests may run it on
``cuda`` in fp32 (TF32 disabled) as the checker.  It is never on the product path.

Pinned against the unmodified reference by ``tests/test_oracle_vs_golden.py``
(golden vectors produced by ``oracle/gen_golden.py`` from ``/root/reference``) and,
when the reference checkout is present, directly by ``tests/test_oracle_vs_reference.py``.

Reference citations (paths relative to the reference checkout):
  * topology         models/models.py:8-145, 213-223, 257-300 (LeakyReLU(0.2) + Dropout)
                     models/vaemodel.py:8-130, 215-230        (ReLU, no Dropout)
  * FC-latent variant models/mymodel.py:51-143 (encoder), :146-230 (decoder), :262-290 (SoftIntroVAE);
                     loop = utils/trainer_fc.py:214-293 (same update rule, noise [B, z_ch], fixed scale)
  * plain-VAE loss   models/lossf.py:5-24
  * introspective    utils/my_trainer.py:38-48 (calc_kl), :62-78 (calc_reconstruction_loss),
    loss + step      :236-325 (E update / D update)
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class NetCfg:
    """Topology knobs that differ between models/models.py and models/vaemodel.py."""

    in_ch: int
    block_setting: Tuple[Tuple[int, int, int], ...]
    slope: float = 0.2          # LeakyReLU(0.2) in models.py:15,19 ; 0.0 (ReLU) in vaemodel.py:14,18
    p_enc_stem: float = 0.35    # models.py:95   (0 => layer absent, vaemodel.py)
    p_dec_stem: float = 0.25    # models.py:122
    p_dec_tail: float = 0.35    # models.py:140

    @staticmethod
    def soft_intro(in_ch, block_setting):
        return NetCfg(in_ch, tuple(tuple(b) for b in block_setting))

    @staticmethod
    def plain_vae(in_ch, block_setting):
        return NetCfg(in_ch, tuple(tuple(b) for b in block_setting), 0.0, 0.0, 0.0, 0.0)


def encoder_plan(cfg: NetCfg) -> List[Tuple[int, int, int]]:
    """(in_ch, out_ch, stride) of every BuildingBlock, models.py:97-102."""
    plan, cin = [], cfg.in_ch
    for c, n, s in cfg.block_setting:
        for i in range(n):
            plan.append((cin, c, s if i == 0 else 1))
            cin = c
    return plan


def decoder_plan(cfg: NetCfg) -> List[Tuple[int, int, int]]:
    """(in_ch, out_ch, stride) of every UpsampleBuildingkBlock, models.py:124-135."""
    bs = cfg.block_setting[::-1]
    plan, cin = [], cfg.block_setting[-1][0]
    for i, (c, n, s) in enumerate(bs):
        nxt = cfg.in_ch if i == len(bs) - 1 else bs[i + 1][0]
        for j in range(n):
            last = j == n - 1
            cout = nxt if last else c
            plan.append((cin, cout, s if last else 1))
            cin = cout
    return plan


class MaskFeed:
    """Source of dropout keep-masks.  ``masks`` is an ordered list of 0/1 float tensors
    (NCDHW, same shape as the tensor being dropped); when exhausted or ``None`` the
    global torch RNG is used exactly as ``nn.Dropout`` would (F.dropout)."""

    def __init__(self, masks: Optional[Sequence[Tensor]] = None, record: bool = False):
        self._it: Optional[Iterator[Tensor]] = iter(masks) if masks is not None else None
        self.record = record
        self.recorded: List[Tensor] = []

    def apply(self, x: Tensor, p: float, training: bool) -> Tensor:
        if not training or p <= 0.0:
            return x
        if self._it is not None:
            m = next(self._it).to(x.dtype)
            return x * m * (1.0 / (1.0 - p))
        if self.record:
            m = (torch.rand_like(x) >= p).to(x.dtype)
            self.recorded.append(m)
            return x * m * (1.0 / (1.0 - p))
        return F.dropout(x, p, True)


def _act(x: Tensor, slope: float) -> Tensor:
    return F.leaky_relu(x, slope) if slope != 0.0 else F.relu(x)


def _bn(sd: Dict[str, Tensor], key: str, x: Tensor, training: bool) -> Tensor:
    """nn.BatchNorm3d defaults (eps 1e-5, momentum 0.1); updates running stats and
    num_batches_tracked in ``sd`` in place when training."""
    if training:
        sd[key + ".num_batches_tracked"] += 1
    return F.batch_norm(
        x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
        training, 0.1, 1e-5,
    )


def _block(sd, pre, x, stride, res, slope, training, upsample):
    """BuildingBlock (models.py:8-43) / UpsampleBuildingkBlock (models.py:46-80)."""
    h = F.conv3d(x, sd[pre + ".block.0.weight"], None, 1, 1)
    h = _act(_bn(sd, pre + ".block.1", h, training), slope)
    if stride != 1:
        h = F.interpolate(h, scale_factor=float(stride), mode="nearest") if upsample else F.avg_pool3d(h, stride)
    h = F.conv3d(h, sd[pre + ".block.4.weight"], None, 1, 1)
    h = _bn(sd, pre + ".block.5", h, training)
    if res:  # residual only when stride == 1; the projection conv never runs
        h = h + x
    return _act(h, slope)


def encoder_features(sd, x, cfg: NetCfg, training: bool, feed: MaskFeed, prefix="encoder") -> Tensor:
    """ResNetEncoder.blocks, models.py:91-104."""
    h = F.conv3d(x, sd[f"{prefix}.blocks.0.0.weight"], sd[f"{prefix}.blocks.0.0.bias"], 1, 1)
    h = _act(_bn(sd, f"{prefix}.blocks.0.1", h, training), cfg.slope)
    h = feed.apply(h, cfg.p_enc_stem, training)
    for i, (cin, cout, s) in enumerate(encoder_plan(cfg)):
        h = _block(sd, f"{prefix}.blocks.{i + 1}.0", h, s, s == 1, cfg.slope, training, upsample=False)
    return h


def encode(sd, x, cfg: NetCfg, training: bool, feed: Optional[MaskFeed] = None, prefix="encoder"):
    """VAEResNetEncoder.forward, models.py:219-223 -> (mu, logvar)."""
    if isinstance(cfg, FcCfg):
        return fc_encode(sd, x, cfg, training, prefix)
    feed = feed or MaskFeed()
    h = encoder_features(sd, x, cfg, training, feed, prefix)
    mu = F.conv3d(h, sd[f"{prefix}.mu.weight"], sd[f"{prefix}.mu.bias"])
    lv = F.conv3d(h, sd[f"{prefix}.var.weight"], sd[f"{prefix}.var.bias"])
    return mu, lv


def decode(sd, z, cfg: NetCfg, training: bool, feed: Optional[MaskFeed] = None, prefix="decoder") -> Tensor:
    """ResNetDecoder.forward, models.py:116-145."""
    if isinstance(cfg, FcCfg):
        return fc_decode(sd, z, cfg, training, prefix)
    feed = feed or MaskFeed()
    h = F.conv3d(z, sd[f"{prefix}.blocks.0.0.weight"], sd[f"{prefix}.blocks.0.0.bias"])
    h = _act(_bn(sd, f"{prefix}.blocks.0.1", h, training), cfg.slope)
    h = feed.apply(h, cfg.p_dec_stem, training)
    plan = decoder_plan(cfg)
    for i, (cin, cout, s) in enumerate(plan):
        h = _block(sd, f"{prefix}.blocks.{i + 1}.0", h, s, s == 1, cfg.slope, training, upsample=True)
    k = len(plan) + 1
    h = F.conv3d(h, sd[f"{prefix}.blocks.{k}.0.weight"], sd[f"{prefix}.blocks.{k}.0.bias"], 1, 1)
    h = F.relu(h)
    return feed.apply(h, cfg.p_dec_tail, training)


@dataclasses.dataclass(frozen=True)
class FcCfg:
    """models/mymodel.py: SoftIntroVAE(first_ch, second_ch, third_ch, forth_ch, z_ch).  ``grid`` is the spatial size
    in front of the Linear heads -- the reference hard-codes 5x6x5 (mymodel.py:125,151,220: 80x96x80 inputs / 16)."""

    first_ch: int
    second_ch: int
    third_ch: int
    forth_ch: int
    z_ch: int
    grid: Tuple[int, int, int] = (5, 6, 5)
    slope: float = 0.2


def _cbl(sd, pre, i, x, training, slope=0.2, act=True):
    """Conv3d(bias=True) -> BatchNorm3d -> [LeakyReLU(0.2)] at Sequential indices i, i+1 (mymodel.py:56-58 etc.)."""
    h = F.conv3d(x, sd[f"{pre}.{i}.weight"], sd[f"{pre}.{i}.bias"], 1, 1)
    h = _bn(sd, f"{pre}.{i + 1}", h, training)
    return F.leaky_relu(h, slope) if act else h


def fc_encode(sd, x, cfg: FcCfg, training: bool, prefix="encoder"):
    """ResNetVAEencoder.forward, mymodel.py:127-143 -> (mu, logvar) [B, z_ch].  block8 exists but never runs."""
    p, s = prefix, cfg.slope
    h = _cbl(sd, f"{p}.block1", 0, x, training, s)
    h = _cbl(sd, f"{p}.block1", 3, h, training, s)
    h = F.avg_pool3d(h, 2)                                                     # :129
    h = _cbl(sd, f"{p}.block2", 0, h, training, s)
    h = _cbl(sd, f"{p}.block2", 3, h, training, s)
    h = F.avg_pool3d(h, 2)                                                     # :131
    h = _cbl(sd, f"{p}.block3", 0, h, training, s)
    h = _cbl(sd, f"{p}.block3", 3, h, training, s)
    h = F.avg_pool3d(h, 2)                                                     # :133
    h = _cbl(sd, f"{p}.block4short", 0, h, training, s)
    r = _cbl(sd, f"{p}.block5", 0, h, training, s)                             # activation INSIDE the branch
    h = F.leaky_relu(h + r, s)                                                 # :136
    h = _cbl(sd, f"{p}.block6", 0, h, training, s)
    h = F.avg_pool3d(h, 2)                                                     # block6[3]
    h = _cbl(sd, f"{p}.block6", 4, h, training, s)
    r = _cbl(sd, f"{p}.block7", 0, h, training, s)
    r = _cbl(sd, f"{p}.block7", 3, r, training, s, act=False)
    h = F.leaky_relu(h + r, s)                                                 # :139
    h = h.reshape(h.shape[0], -1)                                              # NCDHW flatten, :140
    y = F.linear(h, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"])
    mu, lv = y.chunk(2, dim=1)
    return mu, lv


def fc_decode(sd, z, cfg: FcCfg, training: bool, prefix="decoder") -> Tensor:
    """ResNetDecoder.forward, mymodel.py:217-230: [B, z_ch] -> [B,1,16*grid]."""
    p, s = prefix, cfg.slope
    y = F.relu(F.linear(z.reshape(z.shape[0], -1), sd[f"{p}.dfc.0.weight"], sd[f"{p}.dfc.0.bias"]))
    y = y.reshape(y.shape[0], cfg.forth_ch, *cfg.grid)
    r = _cbl(sd, f"{p}.block1", 0, y, training, s)
    r = _cbl(sd, f"{p}.block1", 3, r, training, s, act=False)
    y = F.leaky_relu(y + r, s)                                                 # :221
    y = _cbl(sd, f"{p}.block2u", 0, y, training, s)
    y = F.interpolate(y, scale_factor=2.0, mode="nearest")
    y = _cbl(sd, f"{p}.block2u", 4, y, training, s)
    r = _cbl(sd, f"{p}.block3", 0, y, training, s)
    r = _cbl(sd, f"{p}.block3", 3, r, training, s, act=False)
    y = F.leaky_relu(y + r, s)                                                 # :224
    for blk in ("block4u", "block5u", "block6u"):
        y = _cbl(sd, f"{p}.{blk}", 0, y, training, s)
        y = F.interpolate(y, scale_factor=2.0, mode="nearest")
        y = _cbl(sd, f"{p}.{blk}", 4, y, training, s)
    y = F.conv3d(y, sd[f"{p}.last_block.0.weight"], sd[f"{p}.last_block.0.bias"], 1, 1)
    return F.relu(y)


def reparameterize(mu: Tensor, logvar: Tensor, eps) -> Tensor:
    """models.py:263-271: three separately rounded fp32 ops; ``eps`` is a tensor
    (train, randn_like) or the python float 0.1 (validation)."""
    std = torch.exp(0.5 * logvar)
    return mu + eps * std


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def calc_kl(logvar: Tensor, mu: Tensor, reduce: str = "mean") -> Tensor:
    """utils/my_trainer.py:38-48."""
    b = mu.size(0)
    mu = mu.reshape(b, -1)
    logvar = logvar.reshape(b, -1)
    kl = -0.5 * torch.sum(1 + logvar - mu ** 2 - logvar.exp(), dim=1)
    if reduce == "mean":
        return kl.mean(dim=0)
    if reduce == "sum":
        return kl.sum()
    return kl


def calc_reconstruction_loss(x: Tensor, recon_x: Tensor, loss_type: str = "mse", reduction: str = "None") -> Tensor:
    """utils/my_trainer.py:62-78 (``loss_type`` is ignored there too)."""
    b = x.size(0)
    per = torch.sum(F.mse_loss(x.reshape(b, -1), recon_x.reshape(b, -1), reduction="none"), dim=1)
    return per.mean(dim=0) if reduction == "mean" else per


def mse_loss(out: Tensor, x: Tensor) -> Tensor:
    """models/lossf.py:5-12."""
    return calc_reconstruction_loss(x, out, reduction="mean")


def kld_loss(mu: Tensor, logvar: Tensor) -> Tensor:
    """models/lossf.py:14-18."""
    return calc_kl(logvar, mu, "mean")


def normal_loss(x_hat, mu, logvar, x, msew=1, kldw=10):
    """models/lossf.py:20-24."""
    mse = mse_loss(x_hat, x) * msew
    kld = kld_loss(mu, logvar) * kldw
    return mse + kld, mse, kld


# --------------------------------------------------------------------------------------
# one Soft-IntroVAE training iteration (utils/my_trainer.py:236-325), gradients only
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class StepHyper:
    beta_rec: float = 1.0
    beta_neg: float = 1024.0
    beta_kl: float = 0.75
    gamma_r: float = 1e-8       # my_trainer.py:193
    scale: Optional[float] = None  # my_trainer.py:194: 8/(80*96*80); None => 8/voxels of the input


def split_state(sd: Dict[str, Tensor]):
    """-> (encoder float params, decoder float params, buffers) name lists, in state_dict order."""
    enc, dec, buf = [], [], []
    for k, v in sd.items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            buf.append(k)
        elif k.startswith("encoder."):
            enc.append(k)
        else:
            dec.append(k)
    return enc, dec, buf


def soft_intro_losses_E(sd, cfg, real, noise, eps3, feed, hp: StepHyper):
    """E-update forward, my_trainer.py:248-284.  ``eps3`` = eps tensors for z, z_rec, z_fake.
    Returns (lossE, dict of scalars, z, fake, rec)."""
    scale = hp.scale if hp.scale is not None else 8.0 / float(real[0].numel())
    fake = decode(sd, noise, cfg, True, feed)                                  # :248
    real_mu, real_lv = encode(sd, real, cfg, True, feed)                       # :250
    z = reparameterize(real_mu, real_lv, eps3[0])                              # :251
    rec = decode(sd, z, cfg, True, feed)                                       # :252
    loss_rec = calc_reconstruction_loss(real, rec, reduction="mean")           # :260
    kl_real = calc_kl(real_lv, real_mu, "mean")                                # :261
    rec_mu, rec_lv = encode(sd, rec.detach(), cfg, True, feed)                 # :266
    z_rec = reparameterize(rec_mu, rec_lv, eps3[1])
    rec_rec = decode(sd, z_rec, cfg, True, feed)
    fake_mu, fake_lv = encode(sd, fake.detach(), cfg, True, feed)              # :267
    z_fake = reparameterize(fake_mu, fake_lv, eps3[2])
    rec_fake = decode(sd, z_fake, cfg, True, feed)
    fake_kl_e = calc_kl(fake_lv, fake_mu, "none")                              # :270
    rec_kl_e = calc_kl(rec_lv, rec_mu, "none")                                 # :271
    loss_fake_rec = calc_reconstruction_loss(fake, rec_fake, reduction="none")  # :274
    loss_rec_rec = calc_reconstruction_loss(rec, rec_rec, reduction="none")    # :275 (rec NOT detached, Q13)
    exp_elbo_fake = (-2 * scale * (hp.beta_rec * loss_fake_rec + hp.beta_neg * fake_kl_e)).exp().mean()
    exp_elbo_rec = (-2 * scale * (hp.beta_rec * loss_rec_rec + hp.beta_neg * rec_kl_e)).exp().mean()
    lossE = scale * (hp.beta_rec * loss_rec + hp.beta_kl * kl_real) + 0.5 * (exp_elbo_fake + exp_elbo_rec)
    lossE = lossE * 10                                                         # :284
    terms = dict(loss_rec=loss_rec, kl_real=kl_real, exp_elbo_fake=exp_elbo_fake, exp_elbo_rec=exp_elbo_rec,
                 fake_kl_e=fake_kl_e.mean(), rec_kl_e=rec_kl_e.mean(),
                 loss_fake_rec_e=loss_fake_rec.mean(), loss_rec_rec_e=loss_rec_rec.mean(), lossE=lossE)
    return lossE, terms, z


def soft_intro_losses_D(sd, cfg, real, noise, z, eps2, feed, hp: StepHyper):
    """D-update forward, my_trainer.py:297-321.  ``eps2`` = eps tensors for z_rec, z_fake."""
    scale = hp.scale if hp.scale is not None else 8.0 / float(real[0].numel())
    fake = decode(sd, noise, cfg, True, feed)                                  # :297
    rec = decode(sd, z.detach(), cfg, True, feed)                              # :298
    loss_rec = calc_reconstruction_loss(real, rec, reduction="mean")           # :301
    rec_mu, rec_lv = encode(sd, rec, cfg, True, feed)                          # :304
    z_rec = reparameterize(rec_mu, rec_lv, eps2[0])
    fake_mu, fake_lv = encode(sd, fake, cfg, True, feed)                       # :307
    z_fake = reparameterize(fake_mu, fake_lv, eps2[1])
    rec_rec = decode(sd, z_rec.detach(), cfg, True, feed)                      # :310
    rec_fake = decode(sd, z_fake.detach(), cfg, True, feed)                    # :311
    loss_rec_rec = calc_reconstruction_loss(rec.detach(), rec_rec, reduction="mean")
    loss_fake_rec = calc_reconstruction_loss(fake.detach(), rec_fake, reduction="mean")
    rec_kl = calc_kl(rec_lv, rec_mu, "mean")
    fake_kl = calc_kl(fake_lv, fake_mu, "mean")
    lossD = scale * (hp.beta_rec * loss_rec + 0.5 * hp.beta_kl * (rec_kl + fake_kl)
                     + hp.gamma_r * 0.5 * hp.beta_rec * (loss_rec_rec + loss_fake_rec))
    lossD = lossD * 10                                                         # :321
    terms = dict(loss_rec_d=loss_rec, rec_kl=rec_kl, fake_kl=fake_kl,
                 loss_rec_rec_d=loss_rec_rec, loss_fake_rec_d=loss_fake_rec, lossD=lossD)
    return lossD, terms


def soft_intro_step_grads(sd, cfg, real, noise, eps5, masks=None, hp: Optional[StepHyper] = None,
                          apply_update: Optional[Callable] = None):
    """One full iteration of my_trainer.py:236-325 WITHOUT the optimiser (unless
    ``apply_update(names, grads, phase)`` is given, which is called where
    ``optimizer_e.step()`` / ``optimizer_d.step()`` sit and may modify ``sd`` in place).

    ``sd`` is modified in place only in its BN buffers (as the reference does).
    ``eps5`` = [eps_z, eps_z_rec(E), eps_z_fake(E), eps_z_rec(D), eps_z_fake(D)].
    ``masks`` = ordered dropout keep-masks in the reference's consumption order
    (SURVEY.md appendix B), or None to draw from the torch RNG.
    Returns (terms: dict[str, float], gradsE: dict[name, Tensor], gradsD: dict[name, Tensor]).
    """
    hp = hp or StepHyper()
    feed = MaskFeed(masks)
    enc_names, dec_names, _ = split_state(sd)
    used = dict(sd)

    # ---- E update: encoder trainable, decoder frozen (my_trainer.py:242-245)
    for k in enc_names:
        used[k] = sd[k].detach().clone().requires_grad_(True)
    for k in dec_names:
        used[k] = sd[k].detach()
    lossE, terms, z = soft_intro_losses_E(used, cfg, real, noise, eps5[:3], feed, hp)
    gl = torch.autograd.grad(lossE, [used[k] for k in enc_names], allow_unused=True)
    gradsE = {k: g for k, g in zip(enc_names, gl) if g is not None}
    if apply_update is not None:
        apply_update(enc_names, gradsE, "E")

    # ---- D update: decoder trainable, encoder frozen (my_trainer.py:291-294)
    used = dict(sd)
    for k in dec_names:
        used[k] = sd[k].detach().clone().requires_grad_(True)
    for k in enc_names:
        used[k] = sd[k].detach()
    lossD, terms_d = soft_intro_losses_D(used, cfg, real, noise, z.detach(), eps5[3:], feed, hp)
    gl = torch.autograd.grad(lossD, [used[k] for k in dec_names], allow_unused=True)
    gradsD = {k: g for k, g in zip(dec_names, gl) if g is not None}
    if apply_update is not None:
        apply_update(dec_names, gradsD, "D")

    terms.update(terms_d)
    return {k: float(v.detach()) for k, v in terms.items()}, gradsE, gradsD


def plain_vae_step_grads(sd, cfg, x, eps, msew=1.0, kldw=1.0):
    """One iteration of train_ResNetVAE (my_trainer.py:588-594) without the optimiser:
    forward = vaemodel.py:226-230, loss = lossf.normal_loss."""
    names = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    used = dict(sd)
    for k in names:
        used[k] = sd[k].detach().clone().requires_grad_(True)
    mu, lv = encode(used, x, cfg, True)
    z = reparameterize(mu, lv, eps)
    x_re = decode(used, z, cfg, True)
    loss, mse, kld = normal_loss(x_re, mu, lv, x, msew, kldw)
    gl = torch.autograd.grad(loss, [used[k] for k in names], allow_unused=True)
    grads = {k: g for k, g in zip(names, gl) if g is not None}
    return dict(loss=float(loss), mse=float(mse), kld=float(kld)), grads, x_re.detach()
