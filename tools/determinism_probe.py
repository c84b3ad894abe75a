"""Run-to-run determinism of one E+D iteration of the CUDA path: identical weights, batch, noise, eps and fed dropout masks,
R repeats; every loss term and every parameter gradient must be BIT-identical between repeats (no float atomics, fixed
reduction orders; a difference means a race or an unordered reduction).  Test infrastructure.

    python tools/determinism_probe.py --vol 40 48 40 --batch 4 --repeats 4
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, nargs=3, default=[40, 48, 40])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--repeats", type=int, default=4)
    ap.add_argument("--philox", action="store_true", help="in-kernel Philox dropout (same seed / epoch) instead of fed masks")
    a = ap.parse_args()
    import sivae_b200
    from sivae_b200 import functional as F, trainer as T
    from tools.parity_probe import _gpu_feed
    dev = torch.device("cuda")
    bs = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
    torch.manual_seed(77)
    net = sivae_b200.SoftIntroVAE(64, bs)
    net.apply(T.init_weights_he)
    net.to(dev).train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    b, (d, h, w) = a.batch, a.vol
    lat = (b, 1, d // 8, h // 8, w // 8)
    g = torch.Generator(device=dev).manual_seed(99)
    real = torch.rand(b, 1, d, h, w, device=dev, generator=g)
    noise = torch.randn(lat, device=dev, generator=g)
    eps = [torch.randn(lat, device=dev, generator=g) for _ in range(5)]
    masks = []
    for ch in "DEDEDED" + "DDEEDD":
        if ch == "E":
            masks.append(torch.rand(b, 64, d, h, w, device=dev, generator=g) >= 0.35)
        else:
            masks.append(torch.rand((b, 256) + lat[2:], device=dev, generator=g) >= 0.25)
            masks.append(torch.rand(b, 1, d, h, w, device=dev, generator=g) >= 0.35)
    feed = _gpu_feed(masks)
    runs = []
    for r in range(a.repeats):
        net.load_state_dict(sd0)
        oe = torch.optim.SGD(net.encoder.parameters(), lr=0.0)
        od = torch.optim.SGD(net.decoder.parameters(), lr=0.0)
        if a.philox:
            F.manual_seed(1234)
        else:
            F.dropout_state.mask_feed = iter(feed)
        F.noise_state.eps_feed = iter(eps)
        try:
            terms = T.soft_intro_train_step(net, real, noise, oe, od, T.StepHyper())
        finally:
            F.dropout_state.mask_feed = None
            F.noise_state.eps_feed = None
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        for p in net.parameters():
            p.grad = None
        runs.append(({k: v.detach().clone() for k, v in terms.items()}, grads,
                     {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}))
    bad = 0
    t0, g0, b0 = runs[0]
    for r, (t, gr, bu) in enumerate(runs[1:], 1):
        for k in t0:
            if not torch.equal(t[k], t0[k]):
                bad += 1
                print(f"repeat {r}: term {k} differs: {float(t0[k])!r} vs {float(t[k])!r}")
        for k in g0:
            if not torch.equal(gr[k], g0[k]):
                bad += 1
                diff = (gr[k].double() - g0[k].double())
                print(f"repeat {r}: grad {k} differs: {int((diff != 0).sum())}/{diff.numel()} elements, "
                      f"max |diff| {float(diff.abs().max()):.3e} (|g| max {float(g0[k].abs().max()):.3e})")
        for k in b0:
            if not torch.equal(bu[k], b0[k]):
                bad += 1
                print(f"repeat {r}: buffer {k} differs")
    print(f"determinism: vol {a.vol} batch {a.batch} {'philox' if a.philox else 'fed masks'}: "
          f"{a.repeats} repeats, {bad} differing tensors -> {'DETERMINISTIC' if bad == 0 else 'NOT deterministic'}")


if __name__ == "__main__":
    main()
