"""Two instruments for the loss-curve bias.  Test infrastructure (uses oracle/).

layers : forward activations after every block of encoder / decoder, ours vs fp32 oracle vs control (autocast bf16),
         at the initial weights and at weights taken from the fp32 training trajectory: where does precision go?
bias   : at FIXED weights, K independent draws of (batch, noise, eps, masks); per parameter the MEAN over draws of
         (g_arm - g_fp32) against its standard error: rounding noise averages out as 1/sqrt(K), a systematic error does
         not.  Also the radial component <mean error, W>/|W|^2 per tensor (exactly 0 for a convolution feeding BatchNorm).

    python tools/parity_probe.py layers --vol 40 48 40 --batch 4 --train-steps 0 40
    python tools/parity_probe.py bias --vol 40 48 40 --batch 4 --train-steps 40 --draws 12
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import loss_curve as L  # noqa: E402


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


class Setup:
    def __init__(self, a):
        import sivae_b200
        from sivae_b200 import trainer as T
        from oracle import sivae_oracle as O
        self.O, self.T = O, T
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.dev = dev = torch.device(a.device)
        if a.net == "config1":
            self.in_ch, self.bs = 12, [[12, 1, 2], [24, 1, 2], [32, 2, 2], [48, 2, 2]]
        else:
            self.in_ch, self.bs = 64, [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
        self.batch = a.batch
        self.vol = d, h, w = a.vol
        f = 2 ** len(self.bs) if a.net == "config1" else 8
        self.lat = (a.batch, 1, d // f, h // f, w // f)
        self.cfg = O.NetCfg.soft_intro(self.in_ch, self.bs)
        torch.manual_seed(77)
        self.net = sivae_b200.SoftIntroVAE(self.in_ch, self.bs)
        self.net.apply(T.init_weights_he)
        self.net.to(dev).train()
        self.sd = {k: v.detach().clone() for k, v in self.net.state_dict().items()}
        self.enc_names, self.dec_names, _ = O.split_state(self.sd)
        for k in self.enc_names + self.dec_names:
            self.sd[k] = torch.nn.Parameter(self.sd[k])
        self.opt = {"E": torch.optim.Adam([self.sd[k] for k in self.enc_names], lr=2e-4),
                    "D": torch.optim.Adam([self.sd[k] for k in self.dec_names], lr=2e-4)}
        gen = torch.Generator().manual_seed(1234)
        self.data = L.synthetic_volumes(a.batch * 4, a.vol, gen).to(dev)
        self.g = torch.Generator(device=dev).manual_seed(4321)
        self.step = 0
        self.c_top = self.bs[-1][0]

    def draw(self, step=None):
        step = self.step if step is None else step
        b, (d, h, w), dev, g = self.batch, self.vol, self.dev, self.g
        real = self.data[(step % 4) * b:(step % 4 + 1) * b]
        noise = torch.randn(self.lat, device=dev, generator=g)
        eps = [torch.randn(self.lat, device=dev, generator=g) for _ in range(5)]
        masks = []
        for ch in "DEDEDED" + "DDEEDD":
            if ch == "E":
                masks.append(torch.rand(b, self.in_ch, d, h, w, device=dev, generator=g) >= 0.35)
            else:
                masks.append(torch.rand((b, self.c_top) + self.lat[2:], device=dev, generator=g) >= 0.25)
                masks.append(torch.rand(b, 1, d, h, w, device=dev, generator=g) >= 0.35)
        return real, noise, eps, masks

    def train_oracle_to(self, target):
        O = self.O

        def upd(names, grads, phase):
            for k in names:
                self.sd[k].grad = grads.get(k)
            self.opt[phase].step()
        while self.step < target:
            real, noise, eps, masks = self.draw()
            O.soft_intro_step_grads(self.sd, self.cfg, real, noise, eps, [m.float() for m in masks], O.StepHyper(),
                                    apply_update=upd)
            self.step += 1

    def snapshot(self):
        return {k: v.detach().clone() for k, v in self.sd.items()}


def _gpu_feed(masks):
    out = []
    for m in masks:
        if m.shape[1] == 1:
            out.append(m[:, 0].to(torch.uint8).contiguous())
        else:
            mm = m.permute(0, 2, 3, 4, 1).to(torch.uint8)
            pad = -m.shape[1] % 64
            if pad:
                mm = torch.cat([mm, torch.ones(*mm.shape[:-1], pad, dtype=torch.uint8, device=mm.device)], dim=-1)
            out.append(mm.contiguous())
    return out


def cmd_layers(a):
    from sivae_b200 import functional as F
    S = Setup(a)
    O = S.O
    for target in a.train_steps:
        S.train_oracle_to(target)
        snap = S.snapshot()
        real, noise, eps, masks = S.draw(step=10_000 + target)   # a probe draw (does not advance the training stream much)
        m_enc, m_ds, m_dt = masks[2], masks[0], masks[1]

        def oracle_pass(autocast):
            rec = []
            orig = O._block

            def blk(sd, pre, x, *args, **kw):
                out = orig(sd, pre, x, *args, **kw)
                rec.append((pre, out.detach().float()))
                return out
            O._block = blk
            try:
                sd = {k: v.clone() for k, v in snap.items()}
                ctx = torch.autocast(S.dev.type, dtype=torch.bfloat16) if autocast else torch.autocast(S.dev.type, enabled=False)
                with torch.no_grad(), ctx:
                    mu, lv = O.encode(sd, real, S.cfg, True, O.MaskFeed([m_enc.float()]))
                    z = O.reparameterize(mu.float(), lv.float(), eps[0])
                    x = O.decode(sd, z, S.cfg, True, O.MaskFeed([m_ds.float(), m_dt.float()]))
            finally:
                O._block = orig
            rec += [("mu", mu.float()), ("logvar", lv.float()), ("x_re", x.float())]
            return rec

        ref = oracle_pass(False)
        ctl = oracle_pass(True)
        S.net.load_state_dict(snap)
        S.net.train()
        ours = []
        hooks = []
        from sivae_b200.models import BuildingBlock
        for name, mod in S.net.named_modules():
            if isinstance(mod, BuildingBlock):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, name=name: ours.append((name, o.detach().float()))))
        F.dropout_state.mask_feed = iter(_gpu_feed([m_enc, m_ds, m_dt]))
        F.noise_state.eps_feed = iter([eps[0]])
        try:
            with torch.no_grad():
                mu, lv = S.net.encode(real)
                z = S.net.reparameterize(mu, lv)
                x = S.net.decode(z)
        finally:
            F.dropout_state.mask_feed = None
            F.noise_state.eps_feed = None
            for h_ in hooks:
                h_.remove()
        ours += [("mu", mu), ("logvar", lv), ("x_re", x)]
        print(f"\n=== forward activations, weights after {target} fp32 steps ({a.net}, vol {a.vol}, batch {a.batch}) ===")
        print(f"{'layer':34s} {'ours cos':>10s} {'ours rel':>10s} {'ctl cos':>10s} {'ctl rel':>10s}   |ref| rms")
        for (n0, r), (_, c), (n1, o) in zip(ref, ctl, ours):
            if o.dim() == 5 and o.shape[1:4] == r.shape[2:]:      # NDHWC (padded) -> NCDHW
                o = o[..., :r.shape[1]].permute(0, 4, 1, 2, 3)
            o = o.reshape(r.shape)
            print(f"{n0:34s} {_cos(o, r):10.6f} {_rel(o, r):10.2e} {_cos(c, r):10.6f} {_rel(c, r):10.2e}   "
                  f"{float(r.pow(2).mean().sqrt()):.3e}")
        sys.stdout.flush()


def cmd_bias(a):
    from sivae_b200 import functional as F, trainer as T
    S = Setup(a)
    O = S.O
    S.train_oracle_to(a.train_steps[0])
    snap = S.snapshot()
    names = S.enc_names + S.dec_names
    acc = {arm: {} for arm in ("fp32", "control", "ours")}
    sq = {arm: {} for arm in ("control", "ours")}
    K = a.draws
    for k_ in range(K):
        real, noise, eps, masks = S.draw(step=k_)
        om = [m.float() for m in masks]
        _, gE, gD = O.soft_intro_step_grads({k: v.clone() for k, v in snap.items()}, S.cfg, real, noise, eps, om, O.StepHyper())
        g0 = {**gE, **gD}
        with torch.autocast(S.dev.type, dtype=torch.bfloat16):
            _, gE, gD = O.soft_intro_step_grads({k: v.clone() for k, v in snap.items()}, S.cfg, real, noise, eps, om, O.StepHyper())
        gc = {k: v.float() for k, v in {**gE, **gD}.items()}
        del om
        S.net.load_state_dict(snap)
        S.net.train()
        oe = torch.optim.SGD(S.net.encoder.parameters(), lr=0.0)
        od = torch.optim.SGD(S.net.decoder.parameters(), lr=0.0)
        F.dropout_state.mask_feed = iter(_gpu_feed(masks))
        F.noise_state.eps_feed = iter(eps)
        try:
            T.soft_intro_train_step(S.net, real, noise, oe, od, T.StepHyper())
        finally:
            F.dropout_state.mask_feed = None
            F.noise_state.eps_feed = None
        go = {k: p.grad.detach().clone() for k, p in S.net.named_parameters() if p.grad is not None}
        for p in S.net.parameters():
            p.grad = None
        for k in g0:
            if k not in go:
                continue
            for arm, gr in (("fp32", g0), ("control", gc), ("ours", go)):
                acc[arm][k] = acc[arm].get(k, 0) + gr[k].double()
            for arm, gr in (("control", gc), ("ours", go)):
                e = (gr[k].double() - g0[k].double())
                sq[arm][k] = sq[arm].get(k, 0) + e * e
    print(f"\n=== gradient bias at the weights after {a.train_steps[0]} fp32 steps, {K} draws ({a.net}, vol {a.vol}, batch {a.batch}) ===")
    print("per tensor: |mean err| / |mean g|  (expected from noise alone: rms err / sqrt(K) / |mean g|);  radial = <mean err, W>/(|W| |mean g|)")
    print(f"{'parameter':44s} {'ours bias':>10s} {'(noise)':>9s} {'ratio':>6s} {'radial':>9s} | {'ctl bias':>10s} {'(noise)':>9s} {'ratio':>6s} {'radial':>9s}")
    tot = {"ours": [0.0, 0.0], "control": [0.0, 0.0]}
    for k in names:
        if k not in acc["fp32"] or k.endswith("blocks.0.0.bias") or k == "decoder.blocks.0.0.weight":
            continue
        m0 = acc["fp32"][k] / K
        n0 = float(m0.norm()) + 1e-300
        w = snap[k].double()
        row = f"{k:44s}"
        for arm in ("ours", "control"):
            me = acc[arm][k] / K - m0
            noise = float((sq[arm][k] / K).sum().sqrt()) / (K ** 0.5)
            bias = float(me.norm())
            radial = float((me * w).sum() / (w.norm() + 1e-300)) / n0
            row += f" {bias / n0:10.3e} {noise / n0:9.2e} {bias / (noise + 1e-300):6.2f} {radial:9.2e} |"
            tot[arm][0] += bias ** 2
            tot[arm][1] += noise ** 2
        print(row)
    for arm in ("ours", "control"):
        print(f"{arm}: total |mean err| / expected-from-noise = {(tot[arm][0] / tot[arm][1]) ** 0.5:.3f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["layers", "bias"])
    ap.add_argument("--vol", type=int, nargs=3, default=[40, 48, 40])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--train-steps", type=int, nargs="+", default=[0])
    ap.add_argument("--draws", type=int, default=12)
    ap.add_argument("--net", default="headline", choices=["headline", "config1"])
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    (cmd_layers if a.cmd == "layers" else cmd_bias)(a)


if __name__ == "__main__":
    main()
