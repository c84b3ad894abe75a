"""The three training loops (utils/my_trainer.py train_soft_intro_vae / train_ResNetVAE, utils/trainer_fc.py
train_soft_intro_vae) end to end on CPU with libsivae.so replaced by its executable specification: return values with
the reference's quirks, the files they write, checkpoints that load back."""
import os

import torch

import sivae_b200
from sivae_b200 import trainer as T
from tests.emu import emulated_kernels

torch.set_num_threads(2)


def _loader(n, vol, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(2, 1, *vol, generator=g), torch.zeros(2)) for _ in range(n)]


def _finite(xs):
    return all(x == x and abs(x) != float("inf") for x in xs)


def test_train_soft_intro_vae_loop(tmp_path):
    bs = [[4, 1, 2], [8, 1, 2], [8, 2, 2]]
    net = sivae_b200.SoftIntroVAE(4, bs)
    for m in net.modules():            # the kernel specification takes explicit keep-masks only: run the loop without dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    path = str(tmp_path) + "/"
    with emulated_kernels():
        out = T.train_soft_intro_vae(net, _loader(2, (8, 8, 8), 1), _loader(1, (8, 8, 8), 2), 1, device=torch.device("cpu"),
                                     path=path)
    assert len(out) == 4 and all(len(v) == 2 for v in out) and all(_finite(v) for v in out)   # appended twice (Q10)
    assert open(path + "train_result.csv").read().strip() == "epoch,train_lossE,train_lossD,val_lossE,val_lossD"
    assert os.path.isfile(path + "loss.txt") and os.path.isfile(path + "kl_losses.txt")
    sd = torch.load(path + "prams/S-IntroVAE_3898_epoch0.pth")
    fresh = sivae_b200.SoftIntroVAE(4, bs)
    fresh.load_state_dict(sd, strict=True)
    assert int(sd["encoder.blocks.0.1.num_batches_tracked"]) == 2 * 5      # 5 encoder passes per iteration (Q15)
    assert next(net.parameters()).device.type == "cpu"


def test_trainer_fc_loop(tmp_path):
    net = sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1))
    w0 = net.decoder.dfc[0].weight.detach().clone()
    path = str(tmp_path) + "/"
    with emulated_kernels():
        out = sivae_b200.trainer_fc.train_soft_intro_vae(net, _loader(2, (16, 16, 16), 3), _loader(1, (16, 16, 16), 4),
                                                         epochs=1, device=torch.device("cpu"), path=path)
    assert len(out) == 4 and all(len(v) == 2 for v in out) and all(_finite(v) for v in out)
    assert os.path.isfile(path + "prams/S-IntroVAE_4184_epoch0.pth")     # trainer_fc.py:418
    sd = torch.load(path + "prams/S-IntroVAE_4184_epoch0.pth")
    sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1)).load_state_dict(sd, strict=True)
    assert float((net.decoder.dfc[0].weight.detach() - w0).abs().max()) > 0       # the Linear heads train
    assert int(sd["encoder.block8.1.num_batches_tracked"]) == 0          # block8 never runs (mymodel.py:108-117)
    # 5 encoder passes per training iteration + 5 in the (eval-mode) validation batch, which does not count
    assert int(sd["encoder.block1.1.num_batches_tracked"]) == 2 * 5


def test_train_resnet_vae_loop(tmp_path):
    bs = [[4, 1, 2], [8, 1, 2], [8, 2, 2]]
    net = sivae_b200.vaemodel.ResNetVAE(4, bs)
    path = str(tmp_path) + "/"
    with emulated_kernels():
        tr, va = T.train_ResNetVAE(net, _loader(2, (8, 8, 8), 5), _loader(1, (8, 8, 8), 6), epochs=1, lr=1e-3, mse_w=1,
                                   kl_w=1, device=torch.device("cpu"), path=path)
    assert len(tr) == 1 and len(va) == 1 and _finite(tr) and _finite(va)
    assert os.path.isfile(path + "ResNetVAE_3898epoch0.pth") and os.path.isfile(path + "resnetvae_weight.pth")
    sivae_b200.vaemodel.ResNetVAE(4, bs).load_state_dict(torch.load(path + "resnetvae_weight.pth"), strict=True)
