"""Host-logic tests (CPU): the drop-in modules + autograd wiring, with libsivae.so replaced by its
executable specification in fp32, must reproduce the reference's golden vectors."""
import os

import pytest
import torch

import sivae_b200
from sivae_b200 import functional as F
from sivae_b200 import trainer as T
from tests.emu import emulated_kernels, masks_to_feed

torch.set_num_threads(2)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_product_path_fails_loudly_without_cuda():
    net = sivae_b200.SoftIntroVAE(8, [[8, 1, 2], [16, 1, 2], [16, 2, 2]])
    with pytest.raises(sivae_b200.kernels.SivaeError):
        net(torch.rand(1, 1, 16, 16, 16))
    with pytest.raises(sivae_b200.kernels.SivaeError):
        T.calc_kl(torch.zeros(2, 1, 2, 2, 2), torch.zeros(2, 1, 2, 2, 2))


def test_every_product_entry_fails_loudly_without_cuda():
    """No CPU or PyTorch fallback anywhere on the product path: FC-latent variant, optimiser, retrieval, pipeline."""
    err = sivae_b200.kernels.SivaeError
    fc = sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1))
    with pytest.raises(err):
        fc(torch.rand(1, 1, 16, 16, 16))
    with pytest.raises(err):
        fc.decode(torch.randn(2, 6))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(err):
        sivae_b200.FusedAdam([p], lr=1e-3).step()
    with pytest.raises(err):
        sivae_b200.topk_similar(torch.randn(3, 4), torch.randn(5, 4), k=2)
    with pytest.raises(err):
        sivae_b200.pipeline.preprocess(torch.rand(1, 1, 4, 4, 4))
    with pytest.raises(err):
        sivae_b200.lossf.normal_loss(torch.rand(1, 1, 4, 4, 4), torch.zeros(1, 1, 1, 1, 1), torch.zeros(1, 1, 1, 1, 1),
                                     torch.rand(1, 1, 4, 4, 4))


def test_state_dict_contract(golden_dir):
    g = _load(golden_dir, "sivae_small.pt")
    net = sivae_b200.SoftIntroVAE(g["in_ch"], g["block_setting"])
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["sd0"].keys())
    for k, v in g["sd0"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    net.load_state_dict(g["sd0"], strict=True)
    # headline net: 126 entries, 26 Conv3d, 18 BatchNorm3d
    big = sivae_b200.SoftIntroVAE(64, [[64, 1, 2], [128, 1, 2], [256, 2, 2]])
    assert len(big.state_dict()) == 126
    assert sum(type(m) is torch.nn.Conv3d for m in big.modules()) == 26
    assert sum(type(m) is torch.nn.BatchNorm3d for m in big.modules()) == 18
    assert sum(p.numel() for p in big.encoder.parameters()) == 7124739
    assert sum(p.numel() for p in big.decoder.parameters()) == 7124225


def test_eval_forward_matches_golden(golden_dir):
    g = _load(golden_dir, "sivae_small.pt")
    net = sivae_b200.SoftIntroVAE(g["in_ch"], g["block_setting"])
    net.load_state_dict(g["sd0"])
    net.eval()
    with emulated_kernels(), torch.no_grad():
        mu, lv = net.encode(g["real"])
        z = net.reparameterize(mu, lv, True)
        x_re = net.decode(z)
    for a, b in ((mu, "mu"), (lv, "logvar"), (z, "z"), (x_re, "x_re")):
        torch.testing.assert_close(a, g["eval"][b], rtol=1e-4, atol=1e-5)


def test_soft_intro_step_matches_golden(golden_dir):
    g = _load(golden_dir, "sivae_small.pt")
    st = g["step"]
    net = sivae_b200.SoftIntroVAE(g["in_ch"], g["block_setting"])
    net.load_state_dict(g["sd0"])
    net.train()
    opt_e = torch.optim.SGD(net.encoder.parameters(), lr=0.0)
    opt_d = torch.optim.SGD(net.decoder.parameters(), lr=0.0)
    hp = T.StepHyper(**st["hyper"])
    F.dropout_state.mask_feed = iter(masks_to_feed(st["masks"]))
    F.noise_state.eps_feed = iter(st["eps"])
    try:
        with emulated_kernels():
            terms = T.soft_intro_train_step(net, g["real"], g["noise"], opt_e, opt_d, hp)
    finally:
        F.dropout_state.mask_feed = None
        F.noise_state.eps_feed = None
    for k in ("lossE", "lossD", "loss_rec", "kl_real", "exp_elbo_fake", "exp_elbo_rec", "rec_kl", "fake_kl",
              "loss_rec_d", "loss_rec_rec_d", "loss_fake_rec_d"):
        assert float(terms[k]) == pytest.approx(st["terms"][k], rel=1e-4, abs=1e-30), k
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(grads) == set(st["gradsE"]) | set(st["gradsD"])
    allref = {**st["gradsE"], **st["gradsD"]}
    for k, ref in allref.items():
        if k.endswith("blocks.0.0.bias"):
            # a conv bias feeding train-mode BatchNorm has an exactly-zero gradient; both sides hold
            # rounding noise only, so compare against the scale of the same layer's weight gradient
            wscale = float(allref[k.replace(".bias", ".weight")].abs().max())
            assert float(grads[k].abs().max()) <= 1e-3 * wscale and float(ref.abs().max()) <= 1e-3 * wscale, k
            continue
        torch.testing.assert_close(grads[k], ref, rtol=2e-3, atol=1e-4 * float(ref.abs().max()) + 1e-12, msg=k)
    sd = net.state_dict()
    for k, v in st["buffers_after"].items():
        torch.testing.assert_close(sd[k], v, rtol=1e-4, atol=1e-5, msg=k)
    # the encoder is left frozen
    assert not any(p.requires_grad for p in net.encoder.parameters())
    assert all(p.requires_grad for p in net.decoder.parameters())


def test_plain_vae_step_matches_golden(golden_dir):
    g = _load(golden_dir, "vae_small.pt")
    st = g["step"]
    net = sivae_b200.vaemodel.ResNetVAE(g["in_ch"], g["block_setting"])
    assert list(net.state_dict().keys()) == list(g["sd0"].keys())
    net.load_state_dict(g["sd0"])
    net.train()
    F.noise_state.eps_feed = iter([st["eps"]])
    try:
        with emulated_kernels():
            x_re, mu, lv = net(g["x"])
            loss, mse, kld = sivae_b200.lossf.normal_loss(x_re, mu, lv, g["x"], 1.0, 1.0)
            loss.backward()
    finally:
        F.noise_state.eps_feed = None
    assert float(loss.detach()) == pytest.approx(st["terms"]["loss"], rel=1e-4)
    assert float(mse.detach()) == pytest.approx(st["terms"]["mse"], rel=1e-4)
    assert float(kld.detach()) == pytest.approx(st["terms"]["kld"], rel=1e-4)
    torch.testing.assert_close(x_re, st["x_re"], rtol=1e-3, atol=5e-4)
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(grads) == set(st["grads"])
    for k, ref in st["grads"].items():
        if k.endswith("blocks.0.0.bias"):
            wscale = float(st["grads"][k.replace(".bias", ".weight")].abs().max())
            assert float(grads[k].abs().max()) <= 1e-3 * wscale, k
            continue
        # BN over 4 samples per channel at the 1x2x1 latent makes this net ill-conditioned: fp32
        # summation-order noise is amplified ~1e3x towards the early layers
        torch.testing.assert_close(grads[k], ref, rtol=5e-3, atol=2e-3 * float(ref.abs().max()) + 1e-9, msg=k)


def test_loss_functions_match_known_answers(golden_dir):
    g = _load(golden_dir, "loss_kat.pt")
    mu, lv, x, y = g["mu"], g["logvar"], g["x"], g["y"]
    with emulated_kernels():
        torch.testing.assert_close(T.calc_kl(lv, mu, "mean"), g["kl_mean"])
        torch.testing.assert_close(T.calc_kl(lv, mu, "sum"), g["kl_sum"])
        torch.testing.assert_close(T.calc_kl(lv, mu, "none"), g["kl_none"])
        torch.testing.assert_close(T.calc_reconstruction_loss(x, y, reduction="mean"), g["rec_mean"])
        torch.testing.assert_close(T.calc_reconstruction_loss(x, y, reduction="none"), g["rec_none"])
        torch.testing.assert_close(torch.stack(sivae_b200.lossf.normal_loss(y, mu, lv, x)), g["lossf_normal"])
        assert torch.equal(F.reparameterize(mu, lv, 0.1), g["z_val"])
        assert torch.equal(F.reparameterize(mu, lv, g["eps"]), g["z_train"])


# ---------------------------------------------------------------------------------------------------------------
# FC-latent variant (models/mymodel.py), SURVEY 8f NEXT-1
# ---------------------------------------------------------------------------------------------------------------
def test_fc_state_dict_contract(golden_dir):
    g = _load(golden_dir, "fc_small.pt")
    net = sivae_b200.mymodel.SoftIntroVAE(*g["chans"], g["z_ch"])
    sd = net.state_dict()
    assert list(sd.keys()) == g["keys"]
    for k, v in g["sd0"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    net.load_state_dict(g["sd0"], strict=True)
    assert net.z_ch == g["z_ch"]
    # BASELINE config 2 (600z_main.py:179): Linear(38400, 1200) and Linear(600, 38400)
    big = sivae_b200.mymodel.SoftIntroVAE(32, 64, 128, 256, 600)
    assert tuple(big.encoder.fc.weight.shape) == (1200, 38400)
    assert tuple(big.decoder.dfc[0].weight.shape) == (38400, 600)


def fc_small_net(grid=(1, 1, 1), chans=(4, 4, 8, 8), z_ch=6, seed=31):
    torch.manual_seed(seed)
    net = sivae_b200.mymodel.SoftIntroVAE(*chans, z_ch, latent_grid=grid)
    net.apply(T.init_weights_he)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    return net


def test_fc_step_matches_oracle_small_grid():
    from oracle import sivae_oracle as O
    from tests.test_oracle_vs_golden import bias_in_front_of_bn
    grid, chans, z_ch, B = (1, 1, 1), (4, 4, 8, 8), 6, 2
    net = fc_small_net(grid, chans, z_ch)
    net.train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(8)
    real = torch.rand(B, 1, 16, 16, 16, generator=gen)
    noise = torch.randn(B, z_ch, generator=gen)
    eps = [torch.randn(B, z_ch, generator=gen) for _ in range(5)]
    hp = dict(beta_rec=1.0, beta_neg=1024.0, beta_kl=0.75, gamma_r=1e-8, scale=8.0 / 16 ** 3)
    ref_terms, gE, gD = O.soft_intro_step_grads(sd, O.FcCfg(*chans, z_ch, grid), real, noise, eps, None, O.StepHyper(**hp))

    opt_e = torch.optim.SGD(net.encoder.parameters(), lr=0.0)
    opt_d = torch.optim.SGD(net.decoder.parameters(), lr=0.0)
    F.noise_state.eps_feed = iter(eps)
    try:
        with emulated_kernels():
            terms = T.soft_intro_train_step(net, real, noise, opt_e, opt_d, T.StepHyper(**hp))
    finally:
        F.noise_state.eps_feed = None
    for k in ("lossE", "lossD", "loss_rec", "kl_real", "exp_elbo_fake", "exp_elbo_rec", "rec_kl", "fake_kl",
              "loss_rec_d", "loss_rec_rec_d", "loss_fake_rec_d"):
        assert float(terms[k]) == pytest.approx(ref_terms[k], rel=2e-4, abs=1e-30), k
    allref = {**gE, **gD}
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    absorbed = {k for k in allref if bias_in_front_of_bn(k, allref) and not k.endswith("encoder.block1.0.bias")}
    assert set(grads) == set(allref) - absorbed          # biases in front of a train-mode BN keep grad None
    assert not any(".block8." in k for k in grads)
    for k, got in grads.items():
        ref = allref[k]
        if k == "encoder.block1.0.bias":                  # stem bias: computed, but mathematically zero
            assert float(got.abs().max()) <= 1e-3 * float(allref["encoder.block1.0.weight"].abs().max())
            continue
        torch.testing.assert_close(got, ref, rtol=2e-3, atol=1e-4 * float(ref.abs().max()) + 1e-12, msg=k)
    after = net.state_dict()
    for k, v in sd.items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            torch.testing.assert_close(after[k], v, rtol=1e-5, atol=1e-6, msg=k)


def test_fc_eval_forward_matches_oracle_small_grid():
    from oracle import sivae_oracle as O
    grid, chans, z_ch = (1, 2, 1), (4, 4, 8, 8), 6
    net = fc_small_net(grid, chans, z_ch, seed=32)
    with torch.no_grad():                                  # non-trivial running statistics
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    net.eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    real = torch.rand(2, 1, 16, 32, 16, generator=torch.Generator().manual_seed(9))
    cfg = O.FcCfg(*chans, z_ch, grid)
    mu_r, lv_r = O.encode(sd, real, cfg, False)
    x_r = O.decode(sd, mu_r, cfg, False)
    with emulated_kernels(), torch.no_grad():
        mu, lv = net.encode(real)
        x_re = net.decode(mu)
    torch.testing.assert_close(mu, mu_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lv, lv_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x_re, x_r, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("width", [64, 4])
def test_deferred_bn_updates_match_direct(width):
    """trainer._fork_join issues the second pass of an independent pair on a side stream with its BatchNorm
    running-statistic updates deferred (functional.deferred_bn / apply_deferred_bn).  Two passes through the same
    layers, deferred and applied afterwards, must leave exactly the buffers that two direct passes leave
    (running <- 0.9 running + 0.1 batch composes in call order; num_batches_tracked += 2)."""
    torch.manual_seed(5)
    bs = [[width, 1, 2], [width, 1, 2], [width, 2, 2]]      # width 4: the kernels run zero-padded to 64 channels
    a = sivae_b200.SoftIntroVAE(width, bs)
    a.apply(T.init_weights_he)
    b = sivae_b200.SoftIntroVAE(width, bs)
    b.load_state_dict(a.state_dict())
    a.train(), b.train()
    assert a.two_stream_ok()
    x1, x2 = torch.rand(2, 1, 8, 8, 8), torch.rand(2, 1, 8, 8, 8) * 3.0
    F.noise_state.eps_feed = iter([torch.zeros(2, 1, 1, 1, 1)] * 4)
    for m in list(a.modules()) + list(b.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                              # the kernel specification takes explicit masks only
    try:
        with emulated_kernels(), torch.no_grad():
            ra = [a.forward(x1)[3], a.forward(x2)[3]]
            with F.deferred_bn() as log:
                rb = [b.forward(x1)[3], b.forward(x2)[3]]
            assert len(log) == 2 * 18                              # 18 BatchNorm layers, two passes
            # nothing was touched while deferred
            assert int(b.state_dict()["encoder.blocks.0.1.num_batches_tracked"]) == 0
            F.apply_deferred_bn(log)
    finally:
        F.noise_state.eps_feed = None
    for u, v in zip(ra, rb):
        torch.testing.assert_close(u, v, rtol=0, atol=0)           # train-mode outputs do not depend on the buffers
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k]) and int(sa[k]) in (0, 2), k
        elif "running" in k:
            torch.testing.assert_close(sb[k], sa[k], rtol=2e-5, atol=1e-6, msg=k)


def test_stride1_block_with_channel_change_runs_the_projection_shortcut():
    """BuildingBlock(in != out, stride 1): the reference adds the 1x1 projection conv of the input to the block output
    (models/models.py:28-43).  No shipped block_setting builds such a block, but a drop-in must not refuse
    it: forward and every gradient against the same block written with torch.nn.functional."""
    import torch.nn.functional as TF
    from sivae_b200.models import BuildingBlock
    torch.manual_seed(9)
    blk = BuildingBlock(64, 128, 1)
    blk.train()
    with torch.no_grad():
        for bn in (blk.block[1], blk.block[5]):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
    x = torch.randn(2, 4, 6, 5, 64)                                  # NDHWC
    x.requires_grad_(True)
    with emulated_kernels():
        out = blk(x)
        g = torch.randn_like(out)
        out.backward(g)
    got = {k: p.grad.clone() for k, p in blk.named_parameters() if p.grad is not None}
    gx = x.grad.clone()
    # the same block in torch
    xr = x.detach().clone().requires_grad_(True)
    ps = {k: p.detach().clone().requires_grad_(True) for k, p in blk.named_parameters()}
    xc = xr.permute(0, 4, 1, 2, 3)
    h = TF.conv3d(xc, ps["block.0.weight"], None, 1, 1)
    h = TF.leaky_relu(TF.batch_norm(h, None, None, ps["block.1.weight"], ps["block.1.bias"], True, 0.1, 1e-5), 0.2)
    h = TF.conv3d(h, ps["block.4.weight"], None, 1, 1)
    h = TF.batch_norm(h, None, None, ps["block.5.weight"], ps["block.5.bias"], True, 0.1, 1e-5)
    ref = TF.leaky_relu(h + TF.conv3d(xc, ps["shortcut.weight"], ps["shortcut.bias"]), 0.2).permute(0, 2, 3, 4, 1)
    ref.backward(g)
    torch.testing.assert_close(out.detach(), ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gx, xr.grad, rtol=1e-3, atol=1e-4)
    assert set(got) == {k for k, p in ps.items() if p.grad is not None}
    for k, v in got.items():
        torch.testing.assert_close(v, ps[k].grad, rtol=2e-3, atol=1e-3 * float(ps[k].grad.abs().max()) + 1e-6, msg=k)


def test_fc_variant_deferred_bn_and_bias_match_direct():
    """FC-latent variant (mymodel.py): every convolution carries a bias in front of its train-mode BatchNorm, which the
    fused units move into the running mean.  A deferred pass (side stream of trainer._fork_join) must leave the same
    buffers as a direct one, incl. that bias term, over two passes through the same layers."""
    torch.manual_seed(6)
    a = sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1))
    a.apply(T.init_weights_he)
    b = sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1))
    b.load_state_dict(a.state_dict())
    a.train(), b.train()
    assert a.two_stream_ok()
    x1, x2 = torch.rand(3, 1, 16, 16, 16), torch.rand(3, 1, 16, 16, 16) * 2.0
    F.noise_state.eps_feed = iter([torch.zeros(3, 6)] * 4)
    try:
        with emulated_kernels(), torch.no_grad():
            ra = [a.forward(x1)[3], a.forward(x2)[3]]
            with F.deferred_bn() as log:
                rb = [b.forward(x1)[3], b.forward(x2)[3]]
            assert int(b.state_dict()["encoder.block1.1.num_batches_tracked"]) == 0
            F.apply_deferred_bn(log)
    finally:
        F.noise_state.eps_feed = None
    for u, v in zip(ra, rb):
        torch.testing.assert_close(u, v, rtol=0, atol=0)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k]), k
        elif "running" in k:
            torch.testing.assert_close(sb[k], sa[k], rtol=2e-5, atol=1e-6, msg=k)
    assert float(sa["encoder.block2.1.running_mean"].abs().max()) > 0
