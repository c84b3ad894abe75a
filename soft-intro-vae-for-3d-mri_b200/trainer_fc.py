"""Training loop of the FC-latent variant: drop-in for the reference's ``utils/trainer_fc.py`` (SURVEY 8f NEXT-1).

``train_soft_intro_vae`` keeps the reference's signature, defaults and return value (trainer_fc.py:128-140,:454).
One iteration is exactly ``trainer.soft_intro_train_step`` (the update rule of trainer_fc.py:220-284 is the one of
my_trainer.py:236-325); what differs from ``utils/my_trainer.py`` and is reproduced here:
  * the latent noise is a vector, ``torch.randn(size=(b_size, model.z_ch))`` (:218, :351);
  * ``scale`` is the constant 8/(80*96*80) whatever the input size (:179);
  * validation runs under ``model.eval()`` with *random* eps in every reparameterisation (the model has no
    ``val_flag``), and both validation losses carry the x10 (:382, :402);
  * the two MultiStepLR schedulers are stepped once, after the last epoch (:447-448);
  * the checkpoint is ``prams/S-IntroVAE_4184_epoch{epoch}.pth`` (:418); loss lists are appended twice per epoch
    (:301-302 and :431-434).
Plot / image side effects (``save_image``, ``train_result``) are reporting, not hot path, and are not reproduced.
"""
from __future__ import annotations

import csv
import os
import random
import time

import numpy as np
import torch
import torch.optim as optim

from . import functional as F
from .optim import FusedAdam
from .trainer import (StepHyper, _dist_rank, calc_kl, calc_reconstruction_loss, init_weights_he,  # noqa: F401
                      soft_intro_train_step)

SCALE = 8.0 / (80 * 96 * 80)          # trainer_fc.py:179


@torch.no_grad()
def soft_intro_val_losses(model, real_batch, noise_batch, hp: StepHyper):
    """Validation losses of one batch, utils/trainer_fc.py:352-402."""
    scale = hp.scale if hp.scale is not None else SCALE
    fake = model.decode(noise_batch)
    real_mu, real_logvar = model.encode(real_batch)
    z = model.reparameterize(real_mu, real_logvar)
    rec = model.decode(z)
    loss_rec = calc_reconstruction_loss(real_batch, rec, reduction="mean")
    kl_real = calc_kl(real_logvar, real_mu, reduce="mean")
    rec_mu, rec_logvar, _, rec_rec = model.forward(rec)
    fake_mu, fake_logvar, _, rec_fake = model.forward(fake)
    fake_kl_e = calc_kl(fake_logvar, fake_mu, reduce="none")
    rec_kl_e = calc_kl(rec_logvar, rec_mu, reduce="none")
    loss_fake_rec = calc_reconstruction_loss(fake, rec_fake, reduction="none")
    loss_rec_rec = calc_reconstruction_loss(rec, rec_rec, reduction="none")
    exp_elbo_fake = (-2 * scale * (hp.beta_rec * loss_fake_rec + hp.beta_neg * fake_kl_e)).exp().mean()
    exp_elbo_rec = (-2 * scale * (hp.beta_rec * loss_rec_rec + hp.beta_neg * rec_kl_e)).exp().mean()
    lossE = (scale * (hp.beta_rec * loss_rec + hp.beta_kl * kl_real) + 0.5 * (exp_elbo_fake + exp_elbo_rec)) * 10
    rec_mu, rec_logvar = model.encode(rec)
    z_rec = model.reparameterize(rec_mu, rec_logvar)
    fake_mu, fake_logvar = model.encode(fake)
    z_fake = model.reparameterize(fake_mu, fake_logvar)
    rec_rec = model.decode(z_rec)
    rec_fake = model.decode(z_fake)
    loss_rec_rec = calc_reconstruction_loss(rec, rec_rec, reduction="mean")
    loss_fake_rec = calc_reconstruction_loss(fake, rec_fake, reduction="mean")
    rec_kl = calc_kl(rec_logvar, rec_mu, reduce="mean")
    fake_kl = calc_kl(fake_logvar, fake_mu, reduce="mean")
    lossD = scale * (loss_rec * hp.beta_rec + 0.5 * hp.beta_kl * (rec_kl + fake_kl)
                     + hp.gamma_r * 0.5 * hp.beta_rec * (loss_rec_rec + loss_fake_rec)) * 10
    return dict(lossE=lossE, lossD=lossD, loss_rec=loss_rec, rec_kl=rec_kl)


def train_soft_intro_vae(model=None, train_loader=None, val_loader=None, epochs=500, lr=2e-4,
                         device=torch.device("cpu"), path="./output_SoftIntroVAE/", beta_rec=1.0, beta_neg=1024.0,
                         beta_kl=0.75, pretrained_path=None, reducers=None):
    """Same signature and return value as utils/trainer_fc.py:128-454 (``reducers`` is the only addition: an optional
    ``(reducer_e, reducer_d)`` pair for one-process-per-GPU data parallelism, see parallel.py)."""
    seed = 77
    rank = _dist_rank()
    if rank == 0:                                                     # one writer per job (ranks share ``path``)
        os.makedirs(os.path.join(path, "prams"), exist_ok=True)
        with open(path + "train_result.csv", "w") as f:
            csv.writer(f).writerow(["epoch", "train_lossE", "train_lossD", "val_lossE", "val_lossD"])
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    F.manual_seed(seed)
    model.to(device)                                                  # :159
    if pretrained_path is not None:
        model.load_state_dict(torch.load(pretrained_path, map_location=device), strict=False)
    Adam = FusedAdam if torch.device(device).type == "cuda" else optim.Adam
    optimizer_e = Adam(model.encoder.parameters(), lr=2e-4)           # ``lr`` is ignored, as in the reference (:165-166)
    optimizer_d = Adam(model.decoder.parameters(), lr=2e-4)
    e_scheduler = optim.lr_scheduler.MultiStepLR(optimizer_e, milestones=(350,), gamma=0.1)
    d_scheduler = optim.lr_scheduler.MultiStepLR(optimizer_d, milestones=(350,), gamma=0.1)
    hp = StepHyper(beta_rec, beta_neg, beta_kl, 1e-8, SCALE)
    model.apply(init_weights_he)                                      # after the optional load (:189)
    red_e, red_d = reducers if reducers is not None else (None, None)
    if rank != 0:
        # identical initialisation on every rank (seed 77 above), then rank-distinct latent noise, eps and dropout
        # streams -- DataParallel replicas see different samples AND different noise (main_DataParallel.py:609)
        torch.manual_seed(seed + rank)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(seed + rank)
        F.manual_seed(seed + rank)

    train_lossE_list, train_lossD_list, val_lossE_list, val_lossD_list = [], [], [], []
    train_lossE = train_lossD = val_lossE = val_lossD = 0.0           # never reset per epoch (:192)
    kls_real, kls_fake, kls_rec, rec_errs = [], [], [], []
    start = time.time()
    for epoch in range(epochs):
        model.train()
        ep = dict(kl_real=[], fake_kl=[], rec_kl=[], loss_rec=[])
        for batch, _labels in train_loader:
            b = batch.size(0)
            noise = torch.randn(size=(b, model.z_ch)).to(device)      # :218
            real = batch.to(device, non_blocking=True)
            terms = soft_intro_train_step(model, real, noise, optimizer_e, optimizer_d, hp, red_e, red_d)
            lE, lD = float(terms["lossE"]), float(terms["lossD"])
            if lE != lE or lD != lD:
                raise SystemError                                      # NaN guard, :286-287
            train_lossE += lE
            train_lossD += lD
            for k in ep:
                ep[k].append(float(terms["loss_rec_d" if k == "loss_rec" else k]))
        train_lossE /= max(len(train_loader), 1)
        train_lossD /= max(len(train_loader), 1)
        train_lossE_list.append(train_lossE)
        train_lossD_list.append(train_lossD)

        model.eval()
        for batch, _labels in val_loader:
            noise = torch.randn(size=(batch.size(0), model.z_ch)).to(device)   # :351
            v = soft_intro_val_losses(model, batch.to(device), noise, hp)
            val_lossE += float(v["lossE"])
            val_lossD += float(v["lossD"])
        val_lossE /= max(len(val_loader), 1)
        val_lossD /= max(len(val_loader), 1)
        val_lossE_list.append(val_lossE)
        val_lossD_list.append(val_lossD)
        for lst, key in ((kls_real, "kl_real"), (kls_fake, "fake_kl"), (kls_rec, "rec_kl"), (rec_errs, "loss_rec")):
            lst.append(float(np.mean(ep[key])) if ep[key] else 0.0)

        if rank == 0:
            torch.save(model.to("cpu").state_dict(), path + f"prams/S-IntroVAE_4184_epoch{epoch}.pth")   # :418-421
            model.to(device)
        print(f"Epoch [{epoch + 1}/{epochs}]  train_lossE:{train_lossE:.3f}  train_lossD:{train_lossD:.3f}  "
              f"val_lossE:{val_lossE:.3f}  val_lossD:{val_lossD:.3f}  total:{(time.time() - start) / 60:.1f}min")
        train_lossE_list.append(train_lossE)                           # appended twice (:431-434)
        train_lossD_list.append(train_lossD)
        val_lossE_list.append(val_lossE)
        val_lossD_list.append(val_lossD)
        if rank == 0:
            with open(path + "/loss.txt", "w") as f:
                for name, lst in (("train_lossE", train_lossE_list), ("val_lossE", val_lossE_list),
                                  ("train_lossD", train_lossD_list), ("val_lossD", val_lossD_list)):
                    f.write(name + ":" + ",".join(f"{x:.6f}" for x in lst) + "\n")
            with open(path + "/kl_losses.txt", "w") as f:
                for name, lst in (("kls_real", kls_real), ("kls_fake", kls_fake), ("kls_rec", kls_rec),
                                  ("rec_errs", rec_errs)):
                    f.write(name + ":" + ",".join(f"{x:.6f}" for x in lst) + "\n")
    e_scheduler.step()                                                 # once, after the loop (:447-448)
    d_scheduler.step()
    print("Finished S-IntroVAE Traininig !!")
    model.to("cpu")
    return train_lossE_list, train_lossD_list, val_lossE_list, val_lossD_list
