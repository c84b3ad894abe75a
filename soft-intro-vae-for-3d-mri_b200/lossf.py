"""Drop-in for the reference's ``models/lossf.py`` (plain-VAE loss) on the fused loss kernels.

Same four module-level names, argument order, defaults and return tuples as the reference; the per-sample sums run
as ``functional.mse_persample`` / ``functional.kl_persample`` (csrc/thin_loss.cu), only the batch means and the
weighting stay in torch.
"""
from __future__ import annotations

from . import functional as F


def mse_loss(out, x):
    """mean_b sum_voxels (x - out)^2   (models/lossf.py:5-12)."""
    return F.mse_persample(x, out).mean(dim=0)


def kld_loss(mu, logvar):
    """mean_b -0.5*sum(1 + logvar - mu^2 - exp(logvar))   (models/lossf.py:14-18)."""
    return F.kl_persample(mu, logvar).mean(dim=0)


def _weighted_terms(x_hat, mu, logvar, x, w_rec, w_kl):
    """The two weighted ELBO terms every loss of this file starts from."""
    return mse_loss(x_hat, x) * w_rec, kld_loss(mu, logvar) * w_kl


def normal_loss(x_hat, mu, logvar, x, msew=1, kldw=10):
    """(models/lossf.py:20-24) -> (loss, weighted reconstruction term, weighted KL term)."""
    rec, kl = _weighted_terms(x_hat, mu, logvar, x, msew, kldw)
    return rec + kl, rec, kl


def localized_loss(x_hat, mu, logvar, localize_loss, x, msew=1, kldw=1, localizew=1):
    """(models/lossf.py:26-31; unused by every entry script) -> (loss, reconstruction, KL, localisation term), the
    last one being the batch mean of the row sums of ``localize_loss`` times ``localizew``."""
    rec, kl = _weighted_terms(x_hat, mu, logvar, x, msew, kldw)
    extra = localize_loss.sum(dim=1).mean(dim=0) * localizew
    return rec + kl + extra, rec, kl, extra
