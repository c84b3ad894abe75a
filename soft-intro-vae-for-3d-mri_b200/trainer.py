"""Drop-in for the hot-path part of the reference's ``utils/my_trainer.py``: the introspective loss
functions (``calc_kl`` :38-48, ``calc_reconstruction_loss`` :62-78), one Soft-IntroVAE training
iteration (:236-325) and the ``train_soft_intro_vae`` / ``train_ResNetVAE`` loops (:147-508, :557-652).

To run the *unmodified* reference loop on this implementation instead, pass a
``sivae_b200.SoftIntroVAE`` instance to the reference's ``train_soft_intro_vae`` and rebind
``utils.my_trainer.calc_kl`` / ``calc_reconstruction_loss`` to the functions below (they are looked
up as module globals at call time) -- see INTEGRATION.md.

Hard-coded reference behaviour that is kept: Adam lr 2e-4 for both
optimisers regardless of ``lr``; ``MultiStepLR((350,), 0.1)`` per epoch; gamma_r = 1e-8; seed 77;
train losses x10, validation losses not; every epoch's loss is appended twice; the encoder is left
frozen on return.  The scale ``s`` is 8/(voxels per volume) -- 8/(80*96*80) in the reference.
"""
from __future__ import annotations

import csv
import dataclasses
import os
import random
import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from . import functional as F
from . import lossf
from .optim import FusedAdam


def calc_kl(logvar, mu, reduce="mean"):
    """utils/my_trainer.py:38-48: 'mean' -> scalar batch mean, 'sum' -> scalar, anything else -> [B]."""
    kl = F.kl_persample(mu, logvar)
    if reduce == "mean":
        return kl.mean(dim=0)
    if reduce == "sum":
        return kl.sum()
    return kl


def calc_reconstruction_loss(x, recon_x, loss_type="mse", reduction="None"):
    """utils/my_trainer.py:62-78: per-sample sum of squared errors; 'mean' -> batch mean, else [B].
    ``loss_type`` is accepted and ignored, as in the reference."""
    per = F.mse_persample(x, recon_x)
    return per.mean(dim=0) if reduction == "mean" else per


def init_weights_he(m):
    """utils/my_trainer.py:511-514 (exact-type match on nn.Conv3d / nn.ConvTranspose3d)."""
    if type(m) == nn.Conv3d or type(m) == nn.ConvTranspose3d:
        nn.init.kaiming_normal_(m.weight, nonlinearity="leaky_relu")


def init_weights_he_relu(m):
    """utils/my_trainer.py:516-519."""
    if type(m) == nn.Conv3d or type(m) == nn.ConvTranspose3d:
        nn.init.kaiming_normal_(m.weight, nonlinearity="relu")


@dataclasses.dataclass
class StepHyper:
    beta_rec: float = 1.0
    beta_neg: float = 1024.0
    beta_kl: float = 0.75
    gamma_r: float = 1e-8               # my_trainer.py:193
    scale: Optional[float] = None       # my_trainer.py:194; None -> 8 / voxels-per-volume of the batch


def _dist_rank() -> int:
    """Rank of this process in the default torch.distributed group (0 when not initialised)."""
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _set_requires_grad(module: nn.Module, flag: bool):
    for p in module.parameters():
        p.requires_grad = flag


# Two-stream issue of the independent passes of an iteration (default on; SIVAE_TWO_STREAMS=0 disables): of the 13
# forward passes of utils/my_trainer.py:248-311 these pairs do not depend on each other --
#   E phase: decode(noise) | encode(real) -> z -> decode(z);   forward(rec.detach()) | forward(fake.detach())
#   D phase: decode(noise) | decode(z);   encode(rec) | encode(fake);   decode(z_rec) | decode(z_fake)
# -- and the latent-resolution layers (80-160 CTAs of serial K chains, dozens of 5-10 us BatchNorm launches) leave most
# of the GPU idle.  The second pass of a pair is issued on a side stream; autograd replays each pass's backward on the
# stream its forward ran on, so the backward passes overlap the same way; inside a CUDA-graph capture the pairs become
# parallel branches.  Python still issues the passes in the reference's order, so dropout keys, fed masks and eps
# draws are unchanged, and the side pass defers its BatchNorm running-statistic updates until the streams have joined
# (functional.deferred_bn), which keeps them race-free and in the reference's order.
# Measured (profiles/r02c_ab_streams.md, A/B/A/B on one box, whole-step graph): 64.2 -> 62.5 ms; with the weight-gradient
# stream (functional.wgrad_side_stream) 61.8-62.4 ms.  The gain is bounded by the board's power cap, not by idle SMs: the
# step runs at ~1.3 GHz under sw_power_cap, and overlapping memory-bound passes with tensor-bound ones raises the power
# draw of the same instants.
TWO_STREAMS = os.environ.get("SIVAE_TWO_STREAMS", "1") != "0"
_side_streams = {}
_warned_off = False


def _fork_join(model, device, fn_a, fn_b, allow=True):
    """-> (fn_a(), fn_b()), with fn_b on a side stream when two-stream issue is on and the model allows it.  ``allow`` is
    False with the hook-driven parallel.GradReducer: its buckets are filled from gradient hooks on whichever stream
    the last gradient of a bucket arrives, which must then be the only one."""
    ok = allow and TWO_STREAMS and device.type == "cuda" and hasattr(model, "two_stream_ok")
    if ok:
        if not hasattr(model, "_two_stream_ok"):
            model._two_stream_ok = bool(model.two_stream_ok())
        ok = model._two_stream_ok
    if not ok:
        return fn_a(), fn_b()
    global _warned_off
    if not _warned_off:
        # a parameter's AccumulateGrad node lives on the stream of its first use while the second branch produces
        # gradients on the side stream: the engine orders the two correctly (that is the point); silence its notice
        _warned_off = True
        setter = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if setter is not None:
            setter(False)
    model.prepack()                       # no pack kernel may be launched inside one branch and read by the other
    cur = torch.cuda.current_stream(device)
    side = _side_streams.get(device)
    if side is None:
        side = _side_streams[device] = torch.cuda.Stream(device)
    side.wait_stream(cur)
    ra = fn_a()
    with torch.cuda.stream(side), F.deferred_bn() as log:
        rb = fn_b()
    cur.wait_stream(side)
    # tensors of the side branch that outlive the join (its results, its autograd graph) are only freed after the
    # backward of this phase, i.e. after every reader on either stream has been issued and joined
    F.apply_deferred_bn(log)
    return ra, rb


def _streams_ok(reducer) -> bool:
    """Side streams in forward / backward are compatible with no reducer and with FlatGradReducer (persistent .grad
    views, one exchange after backward), not with the hook-driven GradReducer."""
    return reducer is None or getattr(reducer, "needs_persistent_grads", False)


def _wgrad_stream_ok(reducer) -> bool:
    # the side-stream weight gradients bypass autograd's accumulation (and with it the gradient hooks)
    return F.WGRAD_STREAM and _streams_ok(reducer)


def _run_concurrently(device, fn):
    """Issue ``fn()`` on a dedicated side stream forked from the current one; -> a function that joins it back.
    Used for the encoder's gradient exchange + Adam update, which nothing reads until the D phase encodes."""
    if fn is None:
        return lambda: None
    if device.type != "cuda":
        fn()
        return lambda: None
    cur = torch.cuda.current_stream(device)
    st = _side_streams.get((device, "update"))
    if st is None:
        st = _side_streams[(device, "update")] = torch.cuda.Stream(device)
    st.wait_stream(cur)
    with torch.cuda.stream(st):
        fn()
    return lambda: torch.cuda.current_stream(device).wait_stream(st)


def _encode_sample_decode(model, x):
    mu, logvar = model.encode(x)
    z = model.reparameterize(mu, logvar)
    return mu, logvar, z, model.decode(z)


def _encode_sample(model, x):
    mu, logvar = model.encode(x)
    return mu, logvar, model.reparameterize(mu, logvar)


def _zero_grad(optimizer, reducer):
    # a FlatGradReducer keeps ``.grad`` as views of its flat exchange buffer: clear in place, do not drop them
    optimizer.zero_grad(set_to_none=not getattr(reducer, "needs_persistent_grads", False))


def soft_intro_phase_e(model, real_batch, noise_batch, optimizer_e, hp: Optional[StepHyper] = None, reducer_e=None):
    """Update-E half of one iteration, utils/my_trainer.py:242-287, up to and including ``lossE.backward()``
    (the caller reduces the gradients across ranks, then calls ``optimizer_e.step()``).
    -> (loss terms, z detached for the D phase :298)."""
    hp = hp or StepHyper()
    scale = hp.scale if hp.scale is not None else 8.0 / float(real_batch[0].numel())
    beta_rec, beta_neg, beta_kl = hp.beta_rec, hp.beta_neg, hp.beta_kl
    F.begin_step(real_batch.device)          # new dropout epoch (device-side counter; CUDA-graph safe)
    _set_requires_grad(model.encoder, True)
    _set_requires_grad(model.decoder, False)
    dev = real_batch.device
    ok = _streams_ok(reducer_e)
    fake, (real_mu, real_logvar, z, rec) = _fork_join(
        model, dev, lambda: model.decode(noise_batch), lambda: _encode_sample_decode(model, real_batch), ok)
    (rec_mu, rec_logvar, z_rec, rec_rec), (fake_mu, fake_logvar, z_fake, rec_fake) = _fork_join(
        model, dev, lambda: model.forward(rec.detach()), lambda: model.forward(fake.detach()), ok)
    # per-sample vectors (the mse / kl kernels), then the whole :260-284 assembly in one fused kernel
    r_real = F.mse_persample(real_batch, rec)
    k_real = F.kl_persample(real_mu, real_logvar)
    k_fake = F.kl_persample(fake_mu, fake_logvar)
    k_rec = F.kl_persample(rec_mu, rec_logvar)
    r_fake = F.mse_persample(fake, rec_fake)
    r_rec = F.mse_persample(rec, rec_rec)                                        # rec NOT detached (Q13)
    lossE, loss_rec, lossE_real_kl, exp_elbo_fake, exp_elbo_rec = F.intro_loss_e(
        r_real, k_real, r_fake, k_fake, r_rec, k_rec, scale, beta_rec, beta_kl, beta_neg)
    _zero_grad(optimizer_e, reducer_e)
    with F.wgrad_side_stream(dev, _wgrad_stream_ok(reducer_e)):
        lossE.backward()
    out = dict(lossE=lossE.detach(), loss_rec=loss_rec.detach(), kl_real=lossE_real_kl.detach(),
               exp_elbo_fake=exp_elbo_fake.detach(), exp_elbo_rec=exp_elbo_rec.detach())
    return out, z.detach()


def soft_intro_phase_d(model, real_batch, noise_batch, z, optimizer_d, hp: Optional[StepHyper] = None,
                       reducer_d=None, concurrent=None):
    """Update-D half, utils/my_trainer.py:291-323, up to and including ``lossD.backward()``.  ``concurrent``: optional
    callable issued on a side stream next to the first two decoder passes (they read neither encoder weights nor
    encoder gradients: the caller passes the encoder's gradient exchange + optimiser step) and joined before the
    first encoder pass."""
    hp = hp or StepHyper()
    scale = hp.scale if hp.scale is not None else 8.0 / float(real_batch[0].numel())
    beta_rec, beta_kl, gamma_r = hp.beta_rec, hp.beta_kl, hp.gamma_r
    _set_requires_grad(model.encoder, False)
    _set_requires_grad(model.decoder, True)
    dev = real_batch.device
    ok = _streams_ok(reducer_d)
    join = _run_concurrently(dev, concurrent)
    fake, rec = _fork_join(model, dev, lambda: model.decode(noise_batch), lambda: model.decode(z.detach()), ok)
    join()
    (rec_mu, rec_logvar, z_rec), (fake_mu, fake_logvar, z_fake) = _fork_join(
        model, dev, lambda: _encode_sample(model, rec), lambda: _encode_sample(model, fake), ok)
    rec_rec, rec_fake = _fork_join(model, dev, lambda: model.decode(z_rec.detach()), lambda: model.decode(z_fake.detach()),
                                   ok)
    r_real = F.mse_persample(real_batch, rec)
    r_rec_rec = F.mse_persample(rec.detach(), rec_rec)
    r_fake_rec = F.mse_persample(fake.detach(), rec_fake)
    k_rec = F.kl_persample(rec_mu, rec_logvar)
    k_fake = F.kl_persample(fake_mu, fake_logvar)
    lossD, loss_rec, rec_kl, fake_kl, loss_rec_rec, loss_fake_rec = F.intro_loss_d(
        r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, scale, beta_rec, beta_kl, gamma_r)
    _zero_grad(optimizer_d, reducer_d)
    with F.wgrad_side_stream(dev, _wgrad_stream_ok(reducer_d)):
        lossD.backward()
    return dict(lossD=lossD.detach(), loss_rec_d=loss_rec.detach(), rec_kl=rec_kl.detach(), fake_kl=fake_kl.detach(),
                loss_rec_rec_d=loss_rec_rec.detach(), loss_fake_rec_d=loss_fake_rec.detach())


def soft_intro_train_step(model, real_batch, noise_batch, optimizer_e, optimizer_d, hp: Optional[StepHyper] = None,
                          reducer_e=None, reducer_d=None) -> Dict[str, torch.Tensor]:
    """One iteration of utils/my_trainer.py:236-325 (E update then D update), without host syncs.

    ``reducer_e`` / ``reducer_d`` are optional ``parallel.GradReducer`` / ``FlatGradReducer`` objects: their
    ``finish()`` is called between ``backward()`` and ``optimizer.step()``.  Returns the loss terms as device
    tensors.
    """
    out, z = soft_intro_phase_e(model, real_batch, noise_batch, optimizer_e, hp, reducer_e)

    def update_e():
        if reducer_e is not None:
            reducer_e.finish()
        optimizer_e.step()

    if (TWO_STREAMS and getattr(optimizer_e, "graph_safe", False) and real_batch.device.type == "cuda"
            and _streams_ok(reducer_e) and _streams_ok(reducer_d)):
        # the exchange + Adam(E) (which rewrites the encoder's weight packs in its own kernel) run under decode(noise) /
        # decode(z) of the D phase (my_trainer.py:297-298).  Only with FusedAdam: a torch optimiser bumps the version
        # counters on the host, and the re-pack kernels that triggers would race with its update kernels.
        out.update(soft_intro_phase_d(model, real_batch, noise_batch, z, optimizer_d, hp, reducer_d, concurrent=update_e))
    else:
        update_e()
        out.update(soft_intro_phase_d(model, real_batch, noise_batch, z, optimizer_d, hp, reducer_d))
    if reducer_d is not None:
        reducer_d.finish()
    optimizer_d.step()
    return out


@torch.no_grad()
def soft_intro_val_losses(model, real_batch, noise_batch, hp: Optional[StepHyper] = None):
    """Validation losses of one batch, utils/my_trainer.py:386-434 (eval mode, eps = 0.1, no x10; Q8, Q9)."""
    hp = hp or StepHyper()
    scale = hp.scale if hp.scale is not None else 8.0 / float(real_batch[0].numel())
    fake = model.decode(noise_batch)
    real_mu, real_logvar = model.encode(real_batch)
    z = model.reparameterize(real_mu, real_logvar, True)
    rec = model.decode(z)
    loss_rec = calc_reconstruction_loss(real_batch, rec, reduction="mean")
    kl_real = calc_kl(real_logvar, real_mu, reduce="mean")

    def _fwd(x):
        mu, lv = model.encode(x)
        zz = model.reparameterize(mu, lv)      # model.forward draws eps even in validation (:401-402)
        return mu, lv, zz, model.decode(zz)

    rec_mu, rec_logvar, _, rec_rec = _fwd(rec)
    fake_mu, fake_logvar, _, rec_fake = _fwd(fake)
    fake_kl_e = calc_kl(fake_logvar, fake_mu, reduce="none")
    rec_kl_e = calc_kl(rec_logvar, rec_mu, reduce="none")
    loss_fake_rec = calc_reconstruction_loss(fake, rec_fake, reduction="none")
    loss_rec_rec = calc_reconstruction_loss(rec, rec_rec, reduction="none")
    exp_elbo_fake = (-2 * scale * (hp.beta_rec * loss_fake_rec + hp.beta_neg * fake_kl_e)).exp().mean()
    exp_elbo_rec = (-2 * scale * (hp.beta_rec * loss_rec_rec + hp.beta_neg * rec_kl_e)).exp().mean()
    lossE = scale * (hp.beta_rec * loss_rec + hp.beta_kl * kl_real) + 0.5 * (exp_elbo_fake + exp_elbo_rec)
    rec_mu, rec_logvar = model.encode(rec)
    z_rec = model.reparameterize(rec_mu, rec_logvar, True)
    fake_mu, fake_logvar = model.encode(fake)
    z_fake = model.reparameterize(fake_mu, fake_logvar, True)
    rec_rec = model.decode(z_rec)
    rec_fake = model.decode(z_fake)
    loss_rec_rec = calc_reconstruction_loss(rec, rec_rec, reduction="mean")
    loss_fake_rec = calc_reconstruction_loss(fake, rec_fake, reduction="mean")
    rec_kl = calc_kl(rec_logvar, rec_mu, reduce="mean")
    fake_kl = calc_kl(fake_logvar, fake_mu, reduce="mean")
    lossD = scale * (loss_rec * hp.beta_rec + 0.5 * hp.beta_kl * (rec_kl + fake_kl)
                     + hp.gamma_r * 0.5 * hp.beta_rec * (loss_rec_rec + loss_fake_rec))
    return dict(lossE=lossE, lossD=lossD, loss_rec=loss_rec, rec_kl=rec_kl)


def train_soft_intro_vae(model, train_loader, val_loader, epochs, lr=0.001, device=torch.device("cpu"),
                         path="./output_SoftIntroVAE/", beta_rec=1.0, beta_neg=1024.0, beta_kl=0.75,
                         pretrained_path=None, reducers=None):
    """Same signature and return value as utils/my_trainer.py:147-508 (``reducers`` is the only addition:
    an optional ``(GradReducer_e, GradReducer_d)`` pair for one-process-per-GPU data parallelism).

    Plot / image side effects of the reference (``save_image``, ``train_result``) are reporting, not
    hot path, and are not reproduced; the csv header, per-epoch checkpoint and loss text dumps are.
    """
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
    if pretrained_path is not None:
        model.load_state_dict(torch.load(pretrained_path, map_location=device), strict=False)
    # ``lr`` is ignored, as in the reference (:183-184).  On CUDA the update runs as the fused multi-tensor kernel
    # (optim.FusedAdam: same rule and state layout as torch.optim.Adam, plus the bf16 weight re-pack in the same pass)
    Adam = FusedAdam if torch.device(device).type == "cuda" else optim.Adam
    optimizer_e = Adam(model.encoder.parameters(), lr=2e-4)
    optimizer_d = Adam(model.decoder.parameters(), lr=2e-4)
    e_scheduler = optim.lr_scheduler.MultiStepLR(optimizer_e, milestones=(350,), gamma=0.1)
    d_scheduler = optim.lr_scheduler.MultiStepLR(optimizer_d, milestones=(350,), gamma=0.1)
    hp = StepHyper(beta_rec, beta_neg, beta_kl, 1e-8, None)
    model.apply(init_weights_he)                                    # after the optional load (Q6)
    red_e, red_d = reducers if reducers is not None else (None, None)
    if rank != 0:
        # identical initialisation on every rank (seed 77 above), then rank-distinct latent noise, eps and dropout
        # streams -- DataParallel replicas see different samples AND different noise (main_DataParallel.py:609)
        torch.manual_seed(seed + rank)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(seed + rank)
        F.manual_seed(seed + rank)

    train_lossE_list, train_lossD_list, val_lossE_list, val_lossD_list = [], [], [], []
    train_lossE = train_lossD = val_lossE = val_lossD = 0.0         # never reset per epoch (Q10)
    kls_real, kls_fake, kls_rec, rec_errs = [], [], [], []
    start = time.time()
    for epoch in range(epochs):
        model.train()
        ep = dict(kl_real=[], fake_kl=[], rec_kl=[], loss_rec=[])
        for batch, _labels in train_loader:
            b = batch.size(0)
            real = batch.to(device, non_blocking=True)
            lat = (b, 1) + tuple(s // 8 for s in real.shape[2:])    # (b,1,10,12,10) for 80x96x80
            noise = torch.randn(size=lat).to(device)
            terms = soft_intro_train_step(model, real, noise, optimizer_e, optimizer_d, hp, red_e, red_d)
            lE, lD = float(terms["lossE"]), float(terms["lossD"])
            if lE != lE or lD != lD:
                raise SystemError                                    # NaN guard, :327-328
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
            real = batch.to(device)
            lat = (batch.size(0), 1) + tuple(s // 8 for s in real.shape[2:])
            noise = torch.randn(size=lat).to(device)
            v = soft_intro_val_losses(model, real, noise, hp)
            val_lossE += float(v["lossE"])
            val_lossD += float(v["lossD"])
        val_lossE /= max(len(val_loader), 1)
        val_lossD /= max(len(val_loader), 1)
        val_lossE_list.append(val_lossE)
        val_lossD_list.append(val_lossD)
        kls_real.append(float(np.mean(ep["kl_real"])) if ep["kl_real"] else 0.0)
        kls_fake.append(float(np.mean(ep["fake_kl"])) if ep["fake_kl"] else 0.0)
        kls_rec.append(float(np.mean(ep["rec_kl"])) if ep["rec_kl"] else 0.0)
        rec_errs.append(float(np.mean(ep["loss_rec"])) if ep["loss_rec"] else 0.0)

        # per-epoch checkpoint through the CPU, then back (:476-480; re-allocates parameter storage, Q11)
        if rank == 0:
            torch.save(model.to("cpu").state_dict(), path + f"prams/S-IntroVAE_3898_epoch{epoch}.pth")
            model = model.to(device)
        print(f"Epoch[{epoch + 1}/{epochs}] train_lossE:{train_lossE:.3f}  train_lossD:{train_lossD:.3f}  "
              f"val_lossE:{val_lossE:.3f}  val_lossD:{val_lossD:.3f}  total:{(time.time() - start) / 60:.0f}min")
        train_lossE_list.append(train_lossE)                         # appended twice (Q10)
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
        e_scheduler.step()
        d_scheduler.step()
    print("Finished S-IntroVAE Traininig !!")
    model.to("cpu")
    return train_lossE_list, train_lossD_list, val_lossE_list, val_lossD_list


def train_ResNetVAE(net, train_loader, val_loader, epochs=1, lr=0.001, mse_w=1, kl_w=20,
                    device=torch.device("cpu"), path="./output_ResNetVAE/"):
    """utils/my_trainer.py:557-652: plain VAE loop (train uses (mse_w, kl_w), validation the lossf
    defaults -- SURVEY Q20)."""
    os.makedirs(path, exist_ok=True)
    with open(path + "train_result.csv", "w") as f:
        csv.writer(f).writerow(["epoch", "train_loss", "val_loss"])
    optimizer = optim.Adam(net.parameters(), lr)
    net = net.to(device)
    net = net.apply(init_weights_he_relu)
    train_losses, val_losses = [], []
    for epoch in range(epochs):
        net.train()
        run = 0.0
        for inputs, _labels in train_loader:
            inputs = inputs.to(device)
            optimizer.zero_grad()
            x_re, mu, logvar = net.forward(inputs)
            loss, _mse, _kl = lossf.normal_loss(x_re, mu, logvar, inputs, mse_w, kl_w)
            loss.backward()
            optimizer.step()
            run += loss.item()
        train_losses.append(run / max(len(train_loader), 1))
        net.eval()
        run = 0.0
        with torch.no_grad():
            for inputs, _labels in val_loader:
                inputs = inputs.to(device)
                x_re, mu, logvar = net.forward(inputs)
                loss, _mse, _kl = lossf.normal_loss(x_re, mu, logvar, inputs)
                run += loss.item()
        val_losses.append(run / max(len(val_loader), 1))
        if epoch % 10 == 0:
            torch.save(net.to("cpu").state_dict(), path + f"ResNetVAE_3898epoch{epoch}.pth")
            net = net.to(device)
    if epochs != 0:
        net = net.to("cpu")
        torch.save(net.state_dict(), path + "resnetvae_weight.pth")
    return train_losses, val_losses
