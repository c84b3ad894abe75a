"""Loss-curve parity over many training steps (BASELINE.json north_star: "loss-curve parity to the reference over
200 synthetic steps").

Two arms train the SAME Soft-IntroVAE from the SAME init on the SAME synthetic batches, latent noise,
reparameterisation eps and dropout keep-masks, with Adam(lr 2e-4) as utils/my_trainer.py:183-184:

  ours   : sivae_b200 modules on libsivae.so (bf16 NDHWC activations, tcgen05 convolutions, fused BN/act kernels)
  oracle : oracle/sivae_oracle.py, torch fp32 on the same GPU (TF32 off) -- the restatement pinned to the reference's
           golden vectors in tests/test_oracle_vs_golden.py

  control: the same oracle under ``torch.autocast(bfloat16)`` (what stock PyTorch mixed precision gives a user of the
           reference) -- it measures how far ANY bf16 execution of this recipe drifts from fp32.  The recipe is
           violently sensitive: the first Adam updates move every weight by +-lr regardless of gradient magnitude
           and the KL terms are sums of exp(logvar) (kl_real jumps by 4-6 orders of magnitude at step 1), so two
           roundings of the same trajectory separate quickly.  Parity is therefore stated relative to the control.

and the per-step loss terms are compared.  Test infrastructure (it trains the oracle side by side with the product): used by tests/test_loss_curve.py (short runs, asserted) and from the
command line to write profiles/*_loss_curve.{json,md}:

    python tests/loss_curve.py --steps 200 --vol 40 48 40 --batch 4 --out profiles/r01_loss_curve
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TERMS = ("lossE", "lossD", "loss_rec", "kl_real", "loss_rec_d", "rec_kl", "fake_kl")


def synthetic_volumes(n, vol, gen):
    """Smooth 'brain-like' blobs in [0,1] on a zero background (skull-stripped MRI look)."""
    d, h, w = vol
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, d), torch.linspace(-1, 1, h), torch.linspace(-1, 1, w),
                                indexing="ij")
    out = torch.empty(n, 1, d, h, w)
    for i in range(n):
        c = (torch.rand(3, generator=gen) - 0.5) * 0.3
        r = 0.55 + 0.25 * torch.rand(3, generator=gen)
        body = (((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2) < 1.0
        f = 2.0 + 4.0 * torch.rand(3, generator=gen)
        tex = 0.5 + 0.25 * torch.sin(f[0] * zz * 3.14) * torch.cos(f[1] * yy * 3.14) + 0.2 * torch.sin(f[2] * xx * 3.14)
        out[i, 0] = (tex * body).clamp(0, 1)
    return out


def run(steps=200, vol=(16, 24, 16), batch=2, n_batches=4, in_ch=64, block_setting=((64, 1, 2), (128, 1, 2), (256, 2, 2)),
        lr=2e-4, seed=77, device="cuda", verbose=False, control=False, fc=None, perturb=None, warm_start=0):
    """``fc`` = dict(chans=(c1,c2,c3,c4), z_ch=..., grid=(gd,gh,gw)) selects the FC-latent variant (models/mymodel.py +
    utils/trainer_fc.py: vector noise, no dropout, scale fixed at 8/(80*96*80)); ``vol`` must then be 16 * grid.
    ``perturb`` = (seed, rel): every arm starts from the SAME initial weights multiplied by (1 + rel * N(0,1)) --
    replicas for tools/curve_ensemble.py, which measures how far trajectories of ONE arithmetic spread.
    ``warm_start`` = N: the fp32 oracle alone trains N steps first; every arm then continues from ITS weights, BatchNorm
    buffers and Adam moments (exp_avg, exp_avg_sq, step) for ``steps`` recorded steps.  From a cold start the recipe's
    step-1 transient (kl_real jumps from 6e2 to ~1e9 = sum of exp(logvar), set by the few largest logvar elements) is
    amplified chaotically by any rounding, and its gradient spike stays in Adam's second moment (beta2 = 0.999) for the
    whole run, so 200-step curves mostly measure that one spike (profiles/r02_ensemble_*.md); behind the transient the
    curves measure the arithmetic of the training dynamics."""
    import sivae_b200
    from sivae_b200 import functional as F, trainer as T
    from oracle import sivae_oracle as O

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device(device)
    bs = [list(b) for b in block_setting]
    d, h, w = vol
    if fc is not None:
        assert tuple(vol) == tuple(16 * g_ for g_ in fc["grid"]), "FC-latent variant: vol must be 16 * grid"
        lat = (batch, fc["z_ch"])
        cfg = O.FcCfg(*fc["chans"], fc["z_ch"], tuple(fc["grid"]))
    else:
        lat = (batch, 1, d // 8, h // 8, w // 8)
        cfg = O.NetCfg.soft_intro(in_ch, bs)

    torch.manual_seed(seed)
    net = (sivae_b200.mymodel.SoftIntroVAE(*fc["chans"], fc["z_ch"], latent_grid=tuple(fc["grid"])) if fc is not None
           else sivae_b200.SoftIntroVAE(in_ch, bs))
    net.apply(T.init_weights_he)
    if perturb is not None:
        pg = torch.Generator().manual_seed(int(perturb[0]))
        with torch.no_grad():
            for p_ in net.parameters():
                p_.mul_(1.0 + float(perturb[1]) * torch.randn(p_.shape, generator=pg))
    net.to(dev).train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}          # oracle's private copy
    enc_names, dec_names, _ = O.split_state(sd)
    for k in enc_names + dec_names:
        sd[k] = torch.nn.Parameter(sd[k])
    o_opt = {"E": torch.optim.Adam([sd[k] for k in enc_names], lr=lr),
             "D": torch.optim.Adam([sd[k] for k in dec_names], lr=lr)}

    def apply_update(names, grads, phase):
        for k in names:
            sd[k].grad = grads.get(k)                                            # unused params keep grad None
        o_opt[phase].step()

    if control:
        sd_c = {k: v.detach().clone() for k, v in net.state_dict().items()}
        for k in enc_names + dec_names:
            sd_c[k] = torch.nn.Parameter(sd_c[k])
        c_opt = {"E": torch.optim.Adam([sd_c[k] for k in enc_names], lr=lr),
                 "D": torch.optim.Adam([sd_c[k] for k in dec_names], lr=lr)}

        def apply_update_c(names, grads, phase):
            for k in names:
                g_ = grads.get(k)
                sd_c[k].grad = None if g_ is None else g_.float()
            c_opt[phase].step()

    opt_e = torch.optim.Adam(net.encoder.parameters(), lr=lr)
    opt_d = torch.optim.Adam(net.decoder.parameters(), lr=lr)
    hp, ohp = T.StepHyper(), O.StepHyper()
    if fc is not None:
        hp, ohp = T.StepHyper(scale=8.0 / (80 * 96 * 80)), O.StepHyper(scale=8.0 / (80 * 96 * 80))   # trainer_fc.py:179

    gen = torch.Generator().manual_seed(1234)
    data = synthetic_volumes(batch * n_batches, vol, gen).to(dev)
    g = torch.Generator(device=dev).manual_seed(4321)
    c_top = bs[-1][0]
    # dropout masks in the reference's consumption order: D = [dec-stem, dec-tail], E = [enc-stem]
    order = "DEDEDED" + "DDEEDD"      # passes #1..#7 (E phase) and #8..#13 (D phase); model.forward = E then D

    def draw_masks():
        ms = []
        if fc is not None:
            return ms                                      # mymodel.py has no Dropout
        for ch in order:
            if ch == "E":
                ms.append(torch.rand(batch, in_ch, d, h, w, device=dev, generator=g) >= 0.35)
            else:
                ms.append(torch.rand(batch, c_top, d // 8, h // 8, w // 8, device=dev, generator=g) >= 0.25)
                ms.append(torch.rand(batch, 1, d, h, w, device=dev, generator=g) >= 0.35)
        return ms

    def to_feed(ms):
        out = []
        for m in ms:
            if m.shape[1] == 1:
                out.append(m[:, 0].to(torch.uint8).contiguous())
                continue
            mm = m.permute(0, 2, 3, 4, 1).to(torch.uint8)
            pad = -m.shape[1] % 64                       # the module shells pad channel counts to a multiple of 64
            if pad:
                mm = torch.cat([mm, torch.ones(*mm.shape[:-1], pad, dtype=torch.uint8, device=mm.device)], dim=-1)
            out.append(mm.contiguous())
        return out

    curves = {"ours": {k: [] for k in TERMS}, "oracle": {k: [] for k in TERMS}}
    if control:
        curves["control"] = {k: [] for k in TERMS}
    def adopt(opt_dst, params_dst, opt_src, params_src):
        """Give ``opt_dst`` (over ``params_dst``) deep copies of the Adam state ``opt_src`` holds for ``params_src``."""
        for pd, ps in zip(params_dst, params_src):
            st = opt_src.state.get(ps)
            if st:
                opt_dst.state[pd] = {k_: (v_.detach().clone() if torch.is_tensor(v_) else v_) for k_, v_ in st.items()}

    for step in range(-warm_start, steps):
        real = data[(step % n_batches) * batch:(step % n_batches + 1) * batch]
        noise = torch.randn(lat, device=dev, generator=g)
        eps = [torch.randn(lat, device=dev, generator=g) for _ in range(5)]
        masks = draw_masks()
        if step < 0:                                                        # warm start: the fp32 oracle alone
            omasks = None if fc is not None else [m.float() for m in masks]
            O.soft_intro_step_grads(sd, cfg, real, noise, eps, omasks, ohp, apply_update=apply_update)
            if step == -1:
                with torch.no_grad():
                    net.load_state_dict({k: v.detach() for k, v in sd.items()})
                named = dict(net.named_parameters())
                adopt(opt_e, [named[k] for k in enc_names], o_opt["E"], [sd[k] for k in enc_names])
                adopt(opt_d, [named[k] for k in dec_names], o_opt["D"], [sd[k] for k in dec_names])
                if control:
                    with torch.no_grad():
                        for k in sd_c:
                            sd_c[k].copy_(sd[k])
                    adopt(c_opt["E"], [sd_c[k] for k in enc_names], o_opt["E"], [sd[k] for k in enc_names])
                    adopt(c_opt["D"], [sd_c[k] for k in dec_names], o_opt["D"], [sd[k] for k in dec_names])
            continue
        # ---- ours
        F.dropout_state.mask_feed = iter(to_feed(masks))
        F.noise_state.eps_feed = iter(eps)
        try:
            terms = T.soft_intro_train_step(net, real, noise, opt_e, opt_d, hp)
        finally:
            F.dropout_state.mask_feed = None
            F.noise_state.eps_feed = None
        for k in TERMS:
            curves["ours"][k].append(float(terms[k]))
        # ---- oracle
        omasks = None if fc is not None else [m.float() for m in masks]
        oterms, _, _ = O.soft_intro_step_grads(sd, cfg, real, noise, eps, omasks, ohp, apply_update=apply_update)
        for k in TERMS:
            curves["oracle"][k].append(float(oterms[k]))
        if control:
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                cterms, _, _ = O.soft_intro_step_grads(sd_c, cfg, real, noise, eps, omasks, ohp,
                                                       apply_update=apply_update_c)
            for k in TERMS:
                curves["control"][k].append(float(cterms[k]))
        if verbose and (step % 20 == 0 or step == steps - 1):
            print(f"step {step:4d}  lossE {curves['ours']['lossE'][-1]:.5g} / {curves['oracle']['lossE'][-1]:.5g}   "
                  f"lossD {curves['ours']['lossD'][-1]:.5g} / {curves['oracle']['lossD'][-1]:.5g}   "
                  f"rec {curves['ours']['loss_rec'][-1]:.5g} / {curves['oracle']['loss_rec'][-1]:.5g}", flush=True)
    return curves


def deviations(curves, arm="ours"):
    """Per term: median / 90th percentile / max over steps of |arm - oracle| / |oracle|, and the same for the
    10-step moving averages (what a loss plot shows)."""
    out = {}
    for k in TERMS:
        a, b = curves[arm][k], curves["oracle"][k]
        rel = [abs(x - y) / max(abs(y), 1e-30) for x, y in zip(a, b)]
        win = 10
        sm = []
        for i in range(0, max(len(a) - win + 1, 1)):
            ma, mb = sum(a[i:i + win]) / len(a[i:i + win]), sum(b[i:i + win]) / len(b[i:i + win])
            sm.append(abs(ma - mb) / max(abs(mb), 1e-30))
        srt = sorted(rel)
        out[k] = dict(median=statistics.median(rel), p90=srt[int(0.9 * (len(srt) - 1))], max=max(rel),
                      smooth_max=max(sm), first=rel[0], last=rel[-1],
                      oracle_first=b[0], oracle_last=b[-1], ours_last=a[-1])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--vol", type=int, nargs=3, default=[40, 48, 40])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--n-batches", type=int, default=4)
    ap.add_argument("--out", default=None, help="prefix for <out>.json / <out>.md")
    ap.add_argument("--control", action="store_true", help="also train the oracle under bf16 autocast")
    ap.add_argument("--fc", type=int, nargs=5, default=None, metavar=("C1", "C2", "C3", "C4", "Z"),
                    help="FC-latent variant mymodel.SoftIntroVAE(C1,C2,C3,C4,Z); the latent grid is vol/16")
    ap.add_argument("--warm-start", type=int, default=0,
                    help="the fp32 oracle alone trains this many steps first; every arm continues from its state")
    ap.add_argument("--env", nargs="*", default=[], metavar="K=V",
                    help="libsivae kernel-family toggles for this run (e.g. SIVAE_CONV_KD=0), for bisecting")
    a = ap.parse_args()
    for kv in a.env:
        k_, v_ = kv.split("=", 1)
        os.environ[k_] = v_
    fc = None
    if a.fc is not None:
        fc = dict(chans=tuple(a.fc[:4]), z_ch=a.fc[4], grid=tuple(v // 16 for v in a.vol))
    curves = run(a.steps, tuple(a.vol), a.batch, a.n_batches, verbose=True, control=a.control, fc=fc,
                 warm_start=a.warm_start)
    dev = deviations(curves)
    netname = f"mymodel.SoftIntroVAE{tuple(a.fc)} (FC-latent variant)" if a.fc is not None else "headline net"
    lines = [f"# Loss-curve parity, {a.steps} steps, {netname}, volumes {a.vol}, batch {a.batch}, "
             f"{a.n_batches} synthetic batches cycled, identical init / noise / eps / dropout masks"
             + (f", warm start: every arm continues the fp32 oracle's state after {a.warm_start} steps" if a.warm_start
                else ""), "",
             "ours = libsivae.so (bf16 activations) ; oracle = torch fp32 restatement of the reference on the same GPU", "",
             "| term | oracle first | oracle last | ours last | rel.dev median | p90 | max | 10-step-mean max |",
             "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for k, v in dev.items():
        lines.append(f"| {k} | {v['oracle_first']:.5g} | {v['oracle_last']:.5g} | {v['ours_last']:.5g} | "
                     f"{v['median']:.2e} | {v['p90']:.2e} | {v['max']:.2e} | {v['smooth_max']:.2e} |")
    if a.control:
        cdev = deviations(curves, "control")
        lines += ["", "control = the oracle under torch.autocast(bfloat16) vs the fp32 oracle (same table):", "",
                  "| term | control last | rel.dev median | p90 | max | 10-step-mean max |", "|---|---:|---:|---:|---:|---:|"]
        for k, v in cdev.items():
            lines.append(f"| {k} | {v['ours_last']:.5g} | {v['median']:.2e} | {v['p90']:.2e} | {v['max']:.2e} | "
                         f"{v['smooth_max']:.2e} |")
    print("\n".join(lines))
    if a.out:
        with open(a.out + ".json", "w") as f:
            json.dump(dict(config=vars(a), deviations=dev, curves=curves), f)
        with open(a.out + ".md", "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
