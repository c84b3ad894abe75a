"""Drop-in for the reference's ``models/vaemodel.py``: the same topology as ``models.py`` with
``nn.ReLU`` instead of ``LeakyReLU(0.2)`` and **no Dropout** anywhere (vaemodel.py:14,18,87-91,110-114,
127-130); only ``ResNetCAE`` / ``ResNetVAE`` exist there (BASELINE config 1 uses
``ResNetVAE(12, [[12,1,2],[24,1,2],[32,2,2],[48,2,2]])``, vae_main.py:180)."""
from __future__ import annotations

from . import models as _m


class ResNetEncoder(_m.ResNetEncoder):
    _slope = 0.0
    _p_stem = 0.0
    _block_dropout_attr = False


class VAEResNetEncoder(_m.VAEResNetEncoder):
    _slope = 0.0
    _p_stem = 0.0
    _block_dropout_attr = False


class ResNetDecoder(_m.ResNetDecoder):
    _slope = 0.0
    _p_stem = 0.0
    _p_tail = 0.0
    _block_dropout_attr = False


class ResNetCAE(_m.ResNetCAE):
    _enc, _dec = ResNetEncoder, ResNetDecoder


class ResNetVAE(_m.ResNetVAE):
    _enc, _dec = VAEResNetEncoder, ResNetDecoder
