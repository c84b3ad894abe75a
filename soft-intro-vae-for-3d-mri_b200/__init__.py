"""sivae-b200: B200-native (sm_100a) training hot path of Soft-IntroVAE for 3D MRI.

Drop-in for the reference's ``models/models.py`` / ``models/vaemodel.py`` / ``models/lossf.py`` and the
loss functions + training loop of ``utils/my_trainer.py``; all device work goes through the C ABI of
``libsivae.so`` (include/sivae.h).  The directory name carries hyphens, so import it through the
``sivae_b200`` alias module at the repository root.
"""
from . import kernels, functional, models, vaemodel, mymodel, lossf, trainer, trainer_fc, parallel, graph, optim, retrieval, pipeline  # noqa: F401
from .pipeline import GpuInputPipeline  # noqa: F401
from .retrieval import extract_latents, topk_similar  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .models import SoftIntroVAE, ResNetVAE, ResNetCAE  # noqa: F401
from .trainer import (calc_kl, calc_reconstruction_loss, train_soft_intro_vae, soft_intro_train_step,  # noqa: F401
                      init_weights_he, init_weights_he_relu)

__all__ = ["kernels", "functional", "models", "vaemodel", "mymodel", "trainer_fc", "lossf", "trainer", "parallel", "graph", "optim", "FusedAdam", "retrieval", "extract_latents", "topk_similar", "pipeline", "GpuInputPipeline", "SoftIntroVAE",
           "ResNetVAE", "ResNetCAE", "calc_kl", "calc_reconstruction_loss", "train_soft_intro_vae",
           "soft_intro_train_step", "init_weights_he", "init_weights_he_relu"]
