"""Does the recipe itself survive the L-shape (160x192x160, latent 9600) from a random init on uniform-random volumes?
Prints the loss terms of the first steps for the CUDA path and for the fp32 oracle (same init / data / noise, masks and eps
drawn independently).  Test infrastructure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sivae_b200  # noqa: E402
from sivae_b200 import trainer as T, functional as F  # noqa: E402
from oracle import sivae_oracle as O  # noqa: E402

dev = torch.device("cuda")
bs = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
vol = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (160, 192, 160)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
kind = sys.argv[6] if len(sys.argv) > 6 else "rand"
torch.manual_seed(77)
net = sivae_b200.SoftIntroVAE(64, bs)
net.apply(T.init_weights_he)
net.to(dev).train()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
d, h, w = vol
torch.manual_seed(1234)
if kind == "rand":
    real = torch.rand(B, 1, d, h, w, device=dev)
else:
    from tests.loss_curve import synthetic_volumes
    real = synthetic_volumes(B, vol, torch.Generator().manual_seed(1234)).to(dev)
noise = torch.randn(B, 1, d // 8, h // 8, w // 8, device=dev)
oe, od = sivae_b200.FusedAdam(net.encoder.parameters(), lr=2e-4), sivae_b200.FusedAdam(net.decoder.parameters(), lr=2e-4)
F.manual_seed(1234)
for s in range(steps):
    t = T.soft_intro_train_step(net, real, noise, oe, od)
    print(f"ours   step {s}: " + "  ".join(f"{k} {float(v):.4g}" for k, v in t.items() if k in
          ("lossE", "lossD", "loss_rec", "kl_real", "rec_kl", "fake_kl")), flush=True)
del net, oe, od
torch.cuda.empty_cache()
# fp32 oracle, batch 1 of the same data (memory), torch Adam
cfg = O.NetCfg.soft_intro(64, bs)
enc, dec, _ = O.split_state(sd)
for k in enc + dec:
    sd[k] = torch.nn.Parameter(sd[k])
opt = {"E": torch.optim.Adam([sd[k] for k in enc], lr=2e-4), "D": torch.optim.Adam([sd[k] for k in dec], lr=2e-4)}


def upd(names, grads, phase):
    for k in names:
        sd[k].grad = grads.get(k)
    opt[phase].step()


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for s in range(min(steps, 6)):
    eps = [torch.randn(1, 1, d // 8, h // 8, w // 8, device=dev) for _ in range(5)]
    t, _, _ = O.soft_intro_step_grads(sd, cfg, real[:1], noise[:1], eps, None, O.StepHyper(), apply_update=upd)
    print(f"oracle step {s}: " + "  ".join(f"{k} {v:.4g}" for k, v in t.items() if k in
          ("lossE", "lossD", "loss_rec", "kl_real", "rec_kl", "fake_kl")), flush=True)
