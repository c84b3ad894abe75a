"""Drop-in for the reference's ``models/lossf.py`` (plain-VAE loss) on the fused loss kernels."""
from __future__ import annotations

import torch

from . import functional as F


def mse_loss(out, x):
    """mean_b sum_voxels (x - out)^2   (models/lossf.py:5-12)."""
    return F.mse_persample(x, out).mean(dim=0)


def kld_loss(mu, logvar):
    """mean_b -0.5*sum(1 + logvar - mu^2 - exp(logvar))   (models/lossf.py:14-18)."""
    return F.kl_persample(mu, logvar).mean(dim=0)


def normal_loss(x_hat, mu, logvar, x, msew=1, kldw=10):
    """(models/lossf.py:20-24) -> (loss, mse, kld)."""
    mse = mse_loss(x_hat, x) * msew
    kld = kld_loss(mu, logvar) * kldw
    loss = mse + kld
    return loss, mse, kld


def localized_loss(x_hat, mu, logvar, localize_loss, x, msew=1, kldw=1, localizew=1):
    """(models/lossf.py:26-31; unused by every entry script)."""
    mse = mse_loss(x_hat, x) * msew
    kld = kld_loss(mu, logvar) * kldw
    localize_loss = torch.mean(torch.sum(localize_loss, dim=1), dim=0) * localizew
    loss = mse + kld + localize_loss
    return loss, mse, kld, localize_loss
