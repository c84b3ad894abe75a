"""On-GPU input pipeline (NEXT-3): the per-sample work the reference does on the host in its
DataLoader workers, moved behind the H2D copy so that 8 GPUs are not starved by CPU preprocessing.

Reference order (``BrainDataset.__getitem__``, utils/data_load.py:18-24): augment the raw voxel array, then
``_preprocess``.

* ``preprocess``     -- ``BrainDataset._preprocess`` (utils/data_load.py:25-30): per volume
                        ``clip(v, 0, 4*std(v))`` then min-max normalisation to [0, 1].
* ``random_affine``  -- ``tio.OneOf({tio.RandomAffine(degrees=10): 1.0}, p=0.35)`` (aug-z-1200main.py:106-121):
                        with probability ``p`` a volume is resampled (trilinear, outside = volume minimum) through a
                        random rotation of up to +-``degrees`` about each axis and a random per-axis scaling in
                        ``1 +- scales`` (torchio's defaults: ``scales=0.1``, ``center='image'``,
                        ``default_pad_value='minimum'``, ``image_interpolation='linear'``) about the volume centre.
                        torchio is not vendored by the reference nor installed here ("parity
                        unpinned"): the *resampling* kernel is pinned against ``F.grid_sample``; the *parameter
                        sampling* follows torchio's documented semantics (uniform angles / scales, composition
                        scale -> rotate(x) -> rotate(y) -> rotate(z), isotropic voxel spacing assumed).
* ``GpuInputPipeline`` -- both, as one callable for a training loop.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import kernels as K


def _as4(x: torch.Tensor) -> Tuple[torch.Tensor, tuple]:
    shape = tuple(x.shape)
    if x.dim() == 5:
        assert x.shape[1] == 1, "one-channel volumes expected ([B,1,D,H,W])"
        return x.reshape(x.shape[0], *x.shape[2:]), shape
    assert x.dim() == 4
    return x, shape


def preprocess(x: torch.Tensor, cut_range: float = 4.0) -> torch.Tensor:
    """``BrainDataset._preprocess`` for a device batch ``[B,1,D,H,W]`` / ``[B,D,H,W]`` fp32 -> same shape in [0, 1]."""
    x4, shape = _as4(x.contiguous())
    y, _ = K.preprocess_clip_minmax(x4, cut_range)
    return y.reshape(shape)


def affine_matrices(batch: int, vol: Tuple[int, int, int], degrees: float = 10.0, scales: float = 0.1, p: float = 0.35,
                    generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (mats [B,12] fp32 on the CPU, applied [B] bool).  ``mats[b]`` maps OUTPUT voxel (d,h,w,1) to INPUT voxel
    coordinates: the inverse of  T = C . Rz . Ry . Rx . S . C^-1  (C = translation to the volume centre)."""
    d, h, w = vol
    u = torch.rand(batch, 7, generator=generator, dtype=torch.float64)
    applied = u[:, 0] < p
    ang = (u[:, 1:4] * 2 - 1) * math.radians(degrees)
    sc = 1.0 + (u[:, 4:7] * 2 - 1) * scales
    c = torch.tensor([(d - 1) / 2.0, (h - 1) / 2.0, (w - 1) / 2.0], dtype=torch.float64)
    mats = torch.zeros(batch, 12, dtype=torch.float64)
    for b in range(batch):
        if not bool(applied[b]):
            m = torch.eye(3, dtype=torch.float64)
            t = torch.zeros(3, dtype=torch.float64)
        else:
            ax, ay, az = (float(a) for a in ang[b])
            rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]], dtype=torch.float64)
            ry = torch.tensor([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]], dtype=torch.float64)
            rz = torch.tensor([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]], dtype=torch.float64)
            fwd = rz @ ry @ rx @ torch.diag(sc[b])
            m = torch.linalg.inv(fwd)                       # output -> input
            t = c - m @ c
        mats[b] = torch.cat([m, t[:, None]], 1).reshape(12)
    return mats.float(), applied


def random_affine(x: torch.Tensor, degrees: float = 10.0, scales: float = 0.1, p: float = 0.35,
                  generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Random rotation / scaling of each volume of a device batch with probability ``p`` (see module docstring)."""
    x4, shape = _as4(x.contiguous())
    mats, applied = affine_matrices(x4.shape[0], tuple(x4.shape[1:]), degrees, scales, p, generator)
    if not bool(applied.any()):
        return x
    stats = K.volume_stats(x4)                              # pad value = per-volume minimum
    y = K.affine_resample(x4, mats.to(x4.device, non_blocking=True), None, stats)
    return y.reshape(shape)


class GpuInputPipeline:
    """``raw host/device batch -> augmented (training only) -> preprocessed device batch``."""

    def __init__(self, device, train: bool = True, degrees: float = 10.0, scales: float = 0.1, p: float = 0.35,
                 cut_range: float = 4.0, seed: Optional[int] = None):
        self.device = torch.device(device)
        self.train, self.degrees, self.scales, self.p, self.cut_range = train, degrees, scales, p, cut_range
        self.generator = torch.Generator().manual_seed(seed) if seed is not None else None

    def __call__(self, raw: torch.Tensor) -> torch.Tensor:
        x = raw.to(self.device, dtype=torch.float32, non_blocking=True)
        if x.dim() == 4:
            x = x[:, None]
        if self.train and self.p > 0:
            x = random_affine(x, self.degrees, self.scales, self.p, self.generator)
        return preprocess(x, self.cut_range)
