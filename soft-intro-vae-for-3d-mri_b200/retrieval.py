"""Latent extraction and similarity search (NEXT-4).

The reference's goal is content-based retrieval of 3D MRI volumes through the encoder's 1200-dimensional latent
(README.md:4-11).  What it ships is the extraction loop of ``logistic1.ipynb`` cell 7 --
``net.eval(); _, _, z, _ = net.forward(x)`` -- which runs the decoder for nothing and returns a *random sample*
``z = mu + eps * std``; the comparison of latents is left to host-side sklearn code.  This module provides both halves
on the device:

* ``extract_latents`` -- encoder-only forward in eval mode (BatchNorm uses running statistics, dropout off) through the
  libsivae kernels.  ``mode="mu"`` returns the posterior mean (deterministic; what a retrieval index wants),
  ``mode="z"`` reproduces the reference's sampled latent (``reparameterize`` with ``randn_like`` noise).
* ``topk_similar`` -- for every query latent the k most similar database latents under cosine similarity or L2
  distance (``sivae_similarity_topk``).
"""
from __future__ import annotations

from typing import Iterable, Tuple, Union

import torch

from . import kernels as K


@torch.no_grad()
def extract_latents(model, volumes: Union[torch.Tensor, Iterable], device=None, mode: str = "mu",
                    batch_size: int = 8) -> torch.Tensor:
    """-> [N, latent_dim] fp32 on ``device``.

    ``volumes``: a tensor ``[N,1,D,H,W]`` (split into ``batch_size`` chunks) or an iterable of batches / ``(batch,
    label)`` pairs as the reference's DataLoaders yield.  The model is switched to eval mode for the call and restored.
    """
    if mode not in ("mu", "z"):
        raise ValueError("mode must be 'mu' (posterior mean) or 'z' (sampled, as the reference's notebook)")
    device = torch.device(device) if device is not None else next(model.parameters()).device
    was_training = model.training
    model.eval()
    try:
        if isinstance(volumes, torch.Tensor):
            batches = (volumes[i:i + batch_size] for i in range(0, volumes.shape[0], batch_size))
        else:
            batches = volumes
        out = []
        for b in batches:
            if isinstance(b, (tuple, list)):
                b = b[0]
            x = b.to(device, dtype=torch.float32, non_blocking=True)
            mu, logvar = model.encode(x)
            lat = mu if mode == "mu" else model.reparameterize(mu, logvar)
            out.append(lat.reshape(lat.shape[0], -1).float())
        return torch.cat(out, dim=0) if out else torch.empty(0, 0, device=device)
    finally:
        model.train(was_training)


def topk_similar(queries: torch.Tensor, database: torch.Tensor, k: int = 10,
                 metric: str = "cosine") -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (scores [nq,k], index [nq,k] int32), most similar first.  ``metric``: "cosine" (score = cosine similarity) or
    "l2" (score = negative squared Euclidean distance).  1 <= k <= 32."""
    q = queries.reshape(queries.shape[0], -1).float().contiguous()
    d = database.reshape(database.shape[0], -1).float().contiguous()
    return K.similarity_topk(q, d, int(k), metric)
