"""Generate tests/golden/*.pt from the UNMODIFIED reference (authoring container only).

    python -m oracle.gen_golden            # writes tests/golden/{sivae_small,vae_small,loss_kat,fc_small}.pt

The reference has no golden vectors of its own, so these
files -- outputs of the reference's own modules (models/models.py, models/vaemodel.py,
models/lossf.py) and loss functions (utils/my_trainer.py:38-78) on seeded synthetic
inputs -- are the parity pin for ``oracle/sivae_oracle.py`` and, through it, for the
CUDA path.  Dropout masks and reparameterisation noise are *recorded* from the
reference run (F.dropout is wrapped to expose the mask it draws) so the same
draws can be injected into the oracle and into the CUDA kernels.

The training-iteration fixture follows utils/my_trainer.py:236-325 call for call,
using the reference's model methods and loss functions; the noise shape is the only
thing parametrised (the stock loop hard-codes 10x12x10 for 80x96x80 inputs).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class RecordingDropout:
    """Wraps torch.nn.functional.dropout: same semantics, but the keep-mask is drawn
    explicitly (rand >= p) and recorded."""

    def __init__(self):
        self.masks = []
        self._orig = None

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        m = torch.rand_like(x) >= p
        self.masks.append(m)
        return x * m.to(x.dtype) * (1.0 / (1.0 - p))

    def __enter__(self):
        import torch.nn.functional as F
        self._orig = F.dropout
        F.dropout = self
        return self

    def __exit__(self, *a):
        import torch.nn.functional as F
        F.dropout = self._orig


class RecordingRandnLike:
    def __init__(self):
        self.draws = []
        self._orig = None

    def __call__(self, t, *a, **k):
        e = self._orig(t, *a, **k)
        self.draws.append(e.clone())
        return e

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = self
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


def _clone_sd(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def gen_sivae_small(ref_models, ref_trainer):
    torch.manual_seed(1234)
    in_ch, bs = 8, [[8, 1, 2], [16, 1, 2], [16, 2, 2]]
    D, H, W = 16, 24, 16
    B = 2
    net = ref_models.SoftIntroVAE(in_ch, bs)
    net.apply(ref_trainer.init_weights_he)
    # make BN affine params non-trivial so gamma/beta paths are exercised
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    real = torch.rand(B, 1, D, H, W)
    noise = torch.randn(B, 1, D // 8, H // 8, W // 8)
    out = {"in_ch": in_ch, "block_setting": bs, "real": real, "noise": noise, "sd0": _clone_sd(net.state_dict())}

    # ---- eval-mode forward (validation semantics, eps = 0.1; my_trainer.py:391-393)
    net.eval()
    with torch.no_grad():
        mu, lv = net.encode(real)
        z = net.reparameterize(mu, lv, True)
        x_re = net.decode(z)
    out["eval"] = dict(mu=mu, logvar=lv, z=z, x_re=x_re)

    # ---- one training iteration, my_trainer.py:236-325, without the optimiser step
    beta_rec, beta_neg, beta_kl, gamma_r = 1.0, 1024.0, 0.75, 1e-8
    scale = 8.0 / (D * H * W)
    calc_kl, crl = ref_trainer.calc_kl, ref_trainer.calc_reconstruction_loss
    net.train()
    with RecordingDropout() as rd, RecordingRandnLike() as rr:
        for p in net.encoder.parameters():
            p.requires_grad = True
        for p in net.decoder.parameters():
            p.requires_grad = False
        fake = net.decode(noise)
        real_mu, real_lv = net.encode(real)
        z = net.reparameterize(real_mu, real_lv)
        rec = net.decode(z)
        loss_rec = crl(real, rec, loss_type="mse", reduction="mean")
        kl_real = calc_kl(real_lv, real_mu, reduce="mean")
        rec_mu, rec_lv, z_rec, rec_rec = net.forward(rec.detach())
        fake_mu, fake_lv, z_fake, rec_fake = net.forward(fake.detach())
        fake_kl_e = calc_kl(fake_lv, fake_mu, reduce="none")
        rec_kl_e = calc_kl(rec_lv, rec_mu, reduce="none")
        loss_fake_rec = crl(fake, rec_fake, loss_type="mse", reduction="none")
        loss_rec_rec = crl(rec, rec_rec, loss_type="mse", reduction="none")
        exp_elbo_fake = (-2 * scale * (beta_rec * loss_fake_rec + beta_neg * fake_kl_e)).exp().mean()
        exp_elbo_rec = (-2 * scale * (beta_rec * loss_rec_rec + beta_neg * rec_kl_e)).exp().mean()
        lossE = scale * (beta_rec * loss_rec + beta_kl * kl_real) + 0.5 * (exp_elbo_fake + exp_elbo_rec)
        lossE *= 10
        net.zero_grad()
        lossE.backward()
        gradsE = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        termsE = dict(loss_rec=loss_rec, kl_real=kl_real, exp_elbo_fake=exp_elbo_fake, exp_elbo_rec=exp_elbo_rec,
                      fake_kl_e=fake_kl_e.mean(), rec_kl_e=rec_kl_e.mean(),
                      loss_fake_rec_e=loss_fake_rec.mean(), loss_rec_rec_e=loss_rec_rec.mean(), lossE=lossE)
        fwdE = dict(fake=fake.detach(), real_mu=real_mu.detach(), real_logvar=real_lv.detach(), z=z.detach(),
                    rec=rec.detach(), rec_rec=rec_rec.detach(), rec_fake=rec_fake.detach())

        for p in net.encoder.parameters():
            p.requires_grad = False
        for p in net.decoder.parameters():
            p.requires_grad = True
        net.zero_grad()
        for p in net.parameters():
            p.grad = None
        fake = net.decode(noise)
        rec = net.decode(z.detach())
        loss_rec = crl(real, rec, loss_type="mse", reduction="mean")
        rec_mu, rec_lv = net.encode(rec)
        z_rec = net.reparameterize(rec_mu, rec_lv)
        fake_mu, fake_lv = net.encode(fake)
        z_fake = net.reparameterize(fake_mu, fake_lv)
        rec_rec = net.decode(z_rec.detach())
        rec_fake = net.decode(z_fake.detach())
        loss_rec_rec = crl(rec.detach(), rec_rec, loss_type="mse", reduction="mean")
        loss_fake_rec = crl(fake.detach(), rec_fake, loss_type="mse", reduction="mean")
        rec_kl = calc_kl(rec_lv, rec_mu, reduce="mean")
        fake_kl = calc_kl(fake_lv, fake_mu, reduce="mean")
        lossD = scale * (beta_rec * loss_rec + 0.5 * beta_kl * (rec_kl + fake_kl)
                         + gamma_r * 0.5 * beta_rec * (loss_rec_rec + loss_fake_rec))
        lossD *= 10
        lossD.backward()
        gradsD = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        termsD = dict(loss_rec_d=loss_rec, rec_kl=rec_kl, fake_kl=fake_kl,
                      loss_rec_rec_d=loss_rec_rec, loss_fake_rec_d=loss_fake_rec, lossD=lossD)

    terms = {k: float(v.detach()) for k, v in {**termsE, **termsD}.items()}
    out["step"] = dict(
        terms=terms, gradsE=gradsE, gradsD=gradsD, fwdE=fwdE,
        masks=[m.clone() for m in rd.masks], eps=[e.clone() for e in rr.draws],
        buffers_after={k: v.clone() for k, v in net.state_dict().items()
                       if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
        hyper=dict(beta_rec=beta_rec, beta_neg=beta_neg, beta_kl=beta_kl, gamma_r=gamma_r, scale=scale),
    )
    assert len(rr.draws) == 5 and len(rd.masks) == 5 + 16, (len(rr.draws), len(rd.masks))
    return out


def gen_vae_small(ref_vaemodel, ref_lossf, ref_trainer):
    """BASELINE config 1 in miniature: vaemodel.ResNetVAE + lossf.normal_loss, one step
    (utils/my_trainer.py:588-594; vae_main.py:180 uses (12,[[12,1,2],[24,1,2],[32,2,2],[48,2,2]]))."""
    torch.manual_seed(4321)
    in_ch, bs = 4, [[4, 1, 2], [8, 1, 2], [8, 2, 2], [12, 2, 2]]
    D, H, W = 16, 32, 16
    net = ref_vaemodel.ResNetVAE(in_ch, bs)
    net.apply(ref_trainer.init_weights_he_relu)
    x = torch.rand(2, 1, D, H, W)
    out = {"in_ch": in_ch, "block_setting": bs, "x": x, "sd0": _clone_sd(net.state_dict())}
    net.train()
    with RecordingRandnLike() as rr:
        x_re, mu, lv = net.forward(x)
        loss, mse, kld = ref_lossf.normal_loss(x_re, mu, lv, x, 1.0, 1.0)
        net.zero_grad()
        loss.backward()
    out["step"] = dict(
        terms=dict(loss=float(loss), mse=float(mse), kld=float(kld)),
        grads={k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None},
        x_re=x_re.detach(), mu=mu.detach(), logvar=lv.detach(), eps=rr.draws[0],
        buffers_after={k: v.clone() for k, v in net.state_dict().items()
                       if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
    )
    return out


def gen_loss_kat(ref_lossf, ref_trainer):
    """Known-answer vectors for the loss functions in every reduce mode."""
    torch.manual_seed(99)
    mu = torch.randn(3, 1, 2, 3, 2)
    lv = torch.randn(3, 1, 2, 3, 2) * 0.5
    x = torch.rand(3, 1, 4, 6, 4)
    y = torch.rand(3, 1, 4, 6, 4)
    eps = torch.randn_like(mu)
    return dict(
        mu=mu, logvar=lv, x=x, y=y, eps=eps,
        kl_mean=ref_trainer.calc_kl(lv, mu, reduce="mean"),
        kl_sum=ref_trainer.calc_kl(lv, mu, reduce="sum"),
        kl_none=ref_trainer.calc_kl(lv, mu, reduce="none"),
        rec_mean=ref_trainer.calc_reconstruction_loss(x, y, loss_type="mse", reduction="mean"),
        rec_none=ref_trainer.calc_reconstruction_loss(x, y, loss_type="mse", reduction="none"),
        lossf_mse=ref_lossf.mse_loss(y, x), lossf_kld=ref_lossf.kld_loss(mu, lv),
        lossf_normal=torch.stack(ref_lossf.normal_loss(y, mu, lv, x)),
        lossf_normal_w=torch.stack(ref_lossf.normal_loss(y, mu, lv, x, 1.0, 1.0)),
        z_val=mu + 0.1 * torch.exp(0.5 * lv),              # models.py:268-271 (val_flag=True)
        z_train=mu + eps * torch.exp(0.5 * lv),            # models.py:264-266
    )


def _sub(t):
    """Strided 1/64 subsample of a [B,1,D,H,W] volume batch (keeps the fixture small) + its full fp64 sum."""
    return dict(sub=t.detach()[:, :, ::4, ::4, ::4].clone(), sum=float(t.detach().double().sum()))


def gen_fc_small():
    """FC-latent variant: the reference's models/mymodel.py SoftIntroVAE and one iteration of
    utils/trainer_fc.py:214-293 (without the optimiser step) at the only input size the model accepts
    (80x96x80 -> 5x6x5 before the Linear head, mymodel.py:125), with narrow channels so it runs on CPU in seconds.
    ``real`` / ``noise`` are regenerated from the recorded CPU seeds at test time."""
    import importlib
    ref_my = importlib.import_module("models.mymodel")
    ref_tfc = importlib.import_module("utils.trainer_fc")
    chans, z_ch, B = (4, 4, 8, 8), 16, 2
    torch.manual_seed(2024)
    net = ref_my.SoftIntroVAE(*chans, z_ch)
    net.apply(ref_tfc.init_weights_he)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    g = torch.Generator().manual_seed(555)
    real = torch.rand(B, 1, 80, 96, 80, generator=g)
    noise = torch.randn(B, z_ch, generator=g)
    out = {"chans": chans, "z_ch": z_ch, "data_seed": 555, "batch": B, "sd0": _clone_sd(net.state_dict()),
           "keys": list(net.state_dict().keys()), "real_sum": float(real.double().sum()), "noise": noise}

    net.eval()
    with torch.no_grad():
        mu, lv = net.encode(real)
        x_re = net.decode(mu)
    out["eval"] = dict(mu=mu, logvar=lv, x_re_of_mu=_sub(x_re))

    beta_rec, beta_neg, beta_kl, gamma_r = 1.0, 1024.0, 0.75, 1e-8
    scale = 8.0 / (80 * 96 * 80)                                             # trainer_fc.py:179
    calc_kl, crl = ref_tfc.calc_kl, ref_tfc.calc_reconstruction_loss
    net.train()
    with RecordingRandnLike() as rr:
        for p_ in net.encoder.parameters():
            p_.requires_grad = True
        for p_ in net.decoder.parameters():
            p_.requires_grad = False
        fake = net.decode(noise)
        real_mu, real_lv = net.encode(real)
        z = net.reparameterize(real_mu, real_lv)
        rec = net.decode(z)
        loss_rec = crl(real, rec, loss_type="mse", reduction="mean")
        kl_real = calc_kl(real_lv, real_mu, reduce="mean")
        rec_mu, rec_lv, z_rec, rec_rec = net.forward(rec.detach())
        fake_mu, fake_lv, z_fake, rec_fake = net.forward(fake.detach())
        fake_kl_e = calc_kl(fake_lv, fake_mu, reduce="none")
        rec_kl_e = calc_kl(rec_lv, rec_mu, reduce="none")
        loss_fake_rec = crl(fake, rec_fake, loss_type="mse", reduction="none")
        loss_rec_rec = crl(rec, rec_rec, loss_type="mse", reduction="none")
        exp_elbo_fake = (-2 * scale * (beta_rec * loss_fake_rec + beta_neg * fake_kl_e)).exp().mean()
        exp_elbo_rec = (-2 * scale * (beta_rec * loss_rec_rec + beta_neg * rec_kl_e)).exp().mean()
        lossE = scale * (beta_rec * loss_rec + beta_kl * kl_real) + 0.5 * (exp_elbo_fake + exp_elbo_rec)
        lossE = lossE * 10
        net.zero_grad()
        lossE.backward()
        gradsE = {k: p_.grad.clone() for k, p_ in net.named_parameters() if p_.grad is not None}
        termsE = dict(loss_rec=loss_rec, kl_real=kl_real, exp_elbo_fake=exp_elbo_fake, exp_elbo_rec=exp_elbo_rec,
                      fake_kl_e=fake_kl_e.mean(), rec_kl_e=rec_kl_e.mean(),
                      loss_fake_rec_e=loss_fake_rec.mean(), loss_rec_rec_e=loss_rec_rec.mean(), lossE=lossE)
        fwdE = dict(fake=_sub(fake), real_mu=real_mu.detach(), real_logvar=real_lv.detach(), z=z.detach(),
                    rec=_sub(rec), rec_rec=_sub(rec_rec), rec_fake=_sub(rec_fake))

        for p_ in net.encoder.parameters():
            p_.requires_grad = False
        for p_ in net.decoder.parameters():
            p_.requires_grad = True
        for p_ in net.parameters():
            p_.grad = None
        fake = net.decode(noise)
        rec = net.decode(z.detach())
        loss_rec = crl(real, rec, loss_type="mse", reduction="mean")
        rec_mu, rec_lv = net.encode(rec)
        z_rec = net.reparameterize(rec_mu, rec_lv)
        fake_mu, fake_lv = net.encode(fake)
        z_fake = net.reparameterize(fake_mu, fake_lv)
        rec_rec = net.decode(z_rec.detach())
        rec_fake = net.decode(z_fake.detach())
        loss_rec_rec = crl(rec.detach(), rec_rec, loss_type="mse", reduction="mean")
        loss_fake_rec = crl(fake.detach(), rec_fake, loss_type="mse", reduction="mean")
        rec_kl = calc_kl(rec_lv, rec_mu, reduce="mean")
        fake_kl = calc_kl(fake_lv, fake_mu, reduce="mean")
        lossD = scale * (beta_rec * loss_rec + 0.5 * beta_kl * (rec_kl + fake_kl)
                         + gamma_r * 0.5 * beta_rec * (loss_rec_rec + loss_fake_rec))
        lossD = lossD * 10
        lossD.backward()
        gradsD = {k: p_.grad.clone() for k, p_ in net.named_parameters() if p_.grad is not None}
        termsD = dict(loss_rec_d=loss_rec, rec_kl=rec_kl, fake_kl=fake_kl,
                      loss_rec_rec_d=loss_rec_rec, loss_fake_rec_d=loss_fake_rec, lossD=lossD)
    terms = {k: float(v.detach()) for k, v in {**termsE, **termsD}.items()}
    out["step"] = dict(
        terms=terms, gradsE=gradsE, gradsD=gradsD, fwdE=fwdE, eps=[e.clone() for e in rr.draws],
        buffers_after={k: v.clone() for k, v in net.state_dict().items()
                       if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
        hyper=dict(beta_rec=beta_rec, beta_neg=beta_neg, beta_kl=beta_kl, gamma_r=gamma_r, scale=scale),
    )
    assert len(rr.draws) == 5, len(rr.draws)
    return out


def main():
    ref_models, ref_vaemodel, ref_lossf, ref_trainer = import_reference()
    torch.set_num_threads(1)           # deterministic reduction order for the fixtures
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save(gen_sivae_small(ref_models, ref_trainer), os.path.join(GOLDEN, "sivae_small.pt"))
    torch.save(gen_vae_small(ref_vaemodel, ref_lossf, ref_trainer), os.path.join(GOLDEN, "vae_small.pt"))
    torch.save(gen_loss_kat(ref_lossf, ref_trainer), os.path.join(GOLDEN, "loss_kat.pt"))
    torch.save(gen_fc_small(), os.path.join(GOLDEN, "fc_small.pt"))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
