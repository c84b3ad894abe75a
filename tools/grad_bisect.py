"""Bisect the loss-curve divergence: train the fp32 oracle, and at chosen steps compare
the per-parameter gradients of ONE iteration computed from the oracle's current weights by

  oracle   fp32 torch (reference restatement)
  control  the same under torch.autocast(bfloat16)
  ours     libsivae.so, under several kernel-family toggles

on identical batches / noise / eps / dropout masks.  Prints, per arm, the parameters whose gradient is furthest
from the oracle's (cosine, norm ratio, projection <g,g0>/<g0,g0>).  Test infrastructure (uses oracle/).

    python tools/grad_bisect.py --vol 40 48 40 --batch 4 --at 0 40 120
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import loss_curve as L  # noqa: E402

TOGGLES = {
    "default": {},
    "generic": {"SIVAE_CONV_KD": "0", "SIVAE_UPCONV_FUSED": "0", "SIVAE_WGRAD_KW": "0", "SIVAE_UPWGRAD_TALL": "0",
                "SIVAE_NO_FUSED_STATS": "1", "SIVAE_TO1_TAPWISE": "1"},
}


class _Env:
    def __init__(self, kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def compare(name, g, g0):
    a, b = g.flatten().double(), g0.flatten().double()
    nb = b.norm()
    if nb == 0:
        return None
    cos = float(a @ b / (a.norm() * nb + 1e-300))
    return dict(name=name, cos=cos, ratio=float(a.norm() / nb), proj=float(a @ b / (nb * nb)),
                sign=float(((a > 0) == (b > 0)).double().mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, nargs=3, default=[40, 48, 40])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--n-batches", type=int, default=4)
    ap.add_argument("--at", type=int, nargs="+", default=[0, 40, 120])
    ap.add_argument("--worst", type=int, default=12)
    ap.add_argument("--all", action="store_true", help="print every parameter")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--follow", default="oracle", choices=["oracle", "ours"],
                    help="whose Adam trajectory supplies the weights at which the gradients are compared")
    a = ap.parse_args()

    import sivae_b200
    from sivae_b200 import functional as F, trainer as T
    from oracle import sivae_oracle as O

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device(a.device)
    bs = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
    in_ch, batch = 64, a.batch
    d, h, w = a.vol
    lat = (batch, 1, d // 8, h // 8, w // 8)
    cfg = O.NetCfg.soft_intro(in_ch, bs)
    torch.manual_seed(77)
    net = sivae_b200.SoftIntroVAE(in_ch, bs)
    net.apply(T.init_weights_he)
    net.to(dev).train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    enc_names, dec_names, _ = O.split_state(sd)
    for k in enc_names + dec_names:
        sd[k] = torch.nn.Parameter(sd[k])
    o_opt = {"E": torch.optim.Adam([sd[k] for k in enc_names], lr=2e-4),
             "D": torch.optim.Adam([sd[k] for k in dec_names], lr=2e-4)}

    def apply_update(names, grads, phase):
        for k in names:
            sd[k].grad = grads.get(k)
        o_opt[phase].step()

    gen = torch.Generator().manual_seed(1234)
    data = L.synthetic_volumes(batch * a.n_batches, a.vol, gen).to(dev)
    g = torch.Generator(device=dev).manual_seed(4321)
    order = "DEDEDED" + "DDEEDD"

    def draw_masks():
        ms = []
        for ch in order:
            if ch == "E":
                ms.append(torch.rand(batch, in_ch, d, h, w, device=dev, generator=g) >= 0.35)
            else:
                ms.append(torch.rand(batch, 256, d // 8, h // 8, w // 8, device=dev, generator=g) >= 0.25)
                ms.append(torch.rand(batch, 1, d, h, w, device=dev, generator=g) >= 0.35)
        return ms

    def to_feed(ms):
        out = []
        for m in ms:
            if m.shape[1] == 1:
                out.append(m[:, 0].to(torch.uint8).contiguous())
            else:
                out.append(m.permute(0, 2, 3, 4, 1).to(torch.uint8).contiguous())
        return out

    hp, ohp = T.StepHyper(), O.StepHyper()
    last = max(a.at)
    tnet = topt = None
    if a.follow == "ours":
        import copy
        tnet = copy.deepcopy(net)
        topt = (torch.optim.Adam(tnet.encoder.parameters(), lr=2e-4), torch.optim.Adam(tnet.decoder.parameters(), lr=2e-4))
    for step in range(last + 1):
        real = data[(step % a.n_batches) * batch:(step % a.n_batches + 1) * batch]
        noise = torch.randn(lat, device=dev, generator=g)
        eps = [torch.randn(lat, device=dev, generator=g) for _ in range(5)]
        masks = draw_masks()
        omasks = [m.float() for m in masks]
        if step in a.at:
            snap = {k: v.detach().clone() for k, v in (tnet.state_dict() if tnet is not None else sd).items()}
            t0, gE0, gD0 = O.soft_intro_step_grads({k: v.clone() for k, v in snap.items()}, cfg, real, noise, eps,
                                                   omasks, ohp)
            ref = {**gE0, **gD0}
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                tc, gEc, gDc = O.soft_intro_step_grads({k: v.clone() for k, v in snap.items()}, cfg, real, noise, eps,
                                                       omasks, ohp)
            arms = {"control": ({**gEc, **gDc}, tc)}
            for tname, env in TOGGLES.items():
                with _Env(env):
                    net.load_state_dict(snap)
                    net.train()
                    oe = torch.optim.SGD(net.encoder.parameters(), lr=0.0)
                    od = torch.optim.SGD(net.decoder.parameters(), lr=0.0)
                    F.dropout_state.mask_feed = iter(to_feed(masks))
                    F.noise_state.eps_feed = iter(eps)
                    try:
                        terms = T.soft_intro_train_step(net, real, noise, oe, od, hp)
                    finally:
                        F.dropout_state.mask_feed = None
                        F.noise_state.eps_feed = None
                    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
                    for p in net.parameters():
                        p.grad = None
                    arms["ours/" + tname] = (grads, {k: float(v) for k, v in terms.items()})
            print(f"\n================ step {step} ================", flush=True)
            keys = ("lossE", "lossD", "loss_rec", "kl_real", "loss_rec_d", "rec_kl", "fake_kl", "loss_rec_rec_d",
                    "loss_fake_rec_d")
            print("terms        " + "  ".join(f"{k:>14s}" for k in keys))
            print("oracle       " + "  ".join(f"{t0[k]:14.6g}" for k in keys))
            for an, (_, t) in arms.items():
                print(f"{an:12s} " + "  ".join(f"{abs(t[k] - t0[k]) / max(abs(t0[k]), 1e-30):14.3e}" for k in keys))
            for an, (gr, _) in arms.items():
                rows = []
                for k, g0 in ref.items():
                    if k not in gr:
                        rows.append(dict(name=k, cos=float("nan"), ratio=0.0, proj=0.0, sign=0.0))
                        continue
                    r = compare(k, gr[k].float(), g0)
                    if r is not None:
                        rows.append(r)
                rows_sorted = sorted(rows, key=lambda r: (r["cos"] if r["cos"] == r["cos"] else -2))
                sel = rows if a.all else rows_sorted[:a.worst]
                big = [r for r in rows if r["name"].endswith("weight") and "block" in r["name"]]
                mean_cos = sum(r["cos"] for r in big) / max(len(big), 1)
                mean_proj = sum(r["proj"] for r in big) / max(len(big), 1)
                print(f"\n-- {an}: mean cos over block weights {mean_cos:.5f}, mean projection {mean_proj:.5f}; "
                      f"{'all' if a.all else 'worst'} parameters:")
                for r in sel:
                    print(f"   {r['name']:45s} cos {r['cos']:.5f}  |g|/|g0| {r['ratio']:.4f}  proj {r['proj']:.4f}  "
                          f"sign {r['sign']:.4f}")
            sys.stdout.flush()
        if step < last and tnet is not None:
            F.dropout_state.mask_feed = iter(to_feed(masks))
            F.noise_state.eps_feed = iter(eps)
            try:
                T.soft_intro_train_step(tnet, real, noise, topt[0], topt[1], hp)
            finally:
                F.dropout_state.mask_feed = None
                F.noise_state.eps_feed = None
        elif step < last:
            O.soft_intro_step_grads(sd, cfg, real, noise, eps, omasks, ohp, apply_update=apply_update)


if __name__ == "__main__":
    main()
