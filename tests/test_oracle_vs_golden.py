"""Pin oracle/sivae_oracle.py against golden vectors produced by the unmodified
reference (oracle/gen_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import sivae_oracle as O



@pytest.fixture(autouse=True)
def _single_thread():
    """The fixtures were generated with one CPU thread (oracle/gen_golden.py); other test modules change the global
    thread count at import, and the summation order of conv3d's weight gradient follows it."""
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(old)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_loss_known_answers(golden_dir):
    g = _load(golden_dir, "loss_kat.pt")
    mu, lv, x, y = g["mu"], g["logvar"], g["x"], g["y"]
    assert torch.equal(O.calc_kl(lv, mu, "mean"), g["kl_mean"])
    assert torch.equal(O.calc_kl(lv, mu, "sum"), g["kl_sum"])
    assert torch.equal(O.calc_kl(lv, mu, "none"), g["kl_none"])
    assert torch.equal(O.calc_reconstruction_loss(x, y, reduction="mean"), g["rec_mean"])
    assert torch.equal(O.calc_reconstruction_loss(x, y, reduction="none"), g["rec_none"])
    assert torch.equal(O.mse_loss(y, x), g["lossf_mse"])
    assert torch.equal(O.kld_loss(mu, lv), g["lossf_kld"])
    assert torch.equal(torch.stack(O.normal_loss(y, mu, lv, x)), g["lossf_normal"])
    assert torch.equal(torch.stack(O.normal_loss(y, mu, lv, x, 1.0, 1.0)), g["lossf_normal_w"])
    assert torch.equal(O.reparameterize(mu, lv, 0.1), g["z_val"])          # bit-exact
    assert torch.equal(O.reparameterize(mu, lv, g["eps"]), g["z_train"])   # bit-exact


def test_plans_match_headline_net():
    cfg = O.NetCfg.soft_intro(64, [[64, 1, 2], [128, 1, 2], [256, 2, 2]])
    assert O.encoder_plan(cfg) == [(64, 64, 2), (64, 128, 2), (128, 256, 2), (256, 256, 1)]
    assert O.decoder_plan(cfg) == [(256, 256, 1), (256, 128, 2), (128, 64, 2), (64, 64, 2)]


def test_eval_forward(golden_dir):
    g = _load(golden_dir, "sivae_small.pt")
    cfg = O.NetCfg.soft_intro(g["in_ch"], g["block_setting"])
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    mu, lv = O.encode(sd, g["real"], cfg, False)
    z = O.reparameterize(mu, lv, 0.1)
    x_re = O.decode(sd, z, cfg, False)
    for a, b in ((mu, "mu"), (lv, "logvar"), (z, "z"), (x_re, "x_re")):
        torch.testing.assert_close(a, g["eval"][b], rtol=1e-5, atol=1e-6)


def test_soft_intro_step(golden_dir):
    g = _load(golden_dir, "sivae_small.pt")
    st = g["step"]
    cfg = O.NetCfg.soft_intro(g["in_ch"], g["block_setting"])
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    hp = O.StepHyper(**st["hyper"])
    masks = [m.float() for m in st["masks"]]
    terms, gE, gD = O.soft_intro_step_grads(sd, cfg, g["real"], g["noise"], st["eps"], masks, hp)
    for k, v in st["terms"].items():
        assert terms[k] == pytest.approx(v, rel=2e-5, abs=1e-30), k
    assert set(gE) == set(st["gradsE"]) and set(gD) == set(st["gradsD"])
    # unused parameters keep grad None
    assert not any(".shortcut." in k for k in gE) and "encoder.conv.0.weight" not in gE
    allref = {**st["gradsE"], **st["gradsD"]}
    for k, got in {**gE, **gD}.items():
        ref = allref[k]
        if k.endswith("blocks.0.0.bias"):   # exactly-zero gradient (bias in front of train-mode BN): noise only
            wscale = float(allref[k.replace(".bias", ".weight")].abs().max())
            assert float(got.abs().max()) <= 1e-3 * wscale, k
            continue
        torch.testing.assert_close(got, ref, rtol=5e-4, atol=2e-5 * float(ref.abs().max()) + 1e-12, msg=k)
    for k, v in st["buffers_after"].items():
        torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-6, msg=k)
    # 5 encoder + 8 decoder forwards per step
    assert int(sd["encoder.blocks.0.1.num_batches_tracked"]) == 5
    assert int(sd["decoder.blocks.0.1.num_batches_tracked"]) == 8


def test_plain_vae_step(golden_dir):
    g = _load(golden_dir, "vae_small.pt")
    st = g["step"]
    cfg = O.NetCfg.plain_vae(g["in_ch"], g["block_setting"])
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    terms, grads, x_re = O.plain_vae_step_grads(sd, cfg, g["x"], st["eps"], 1.0, 1.0)
    for k, v in st["terms"].items():
        assert terms[k] == pytest.approx(v, rel=2e-5), k
    torch.testing.assert_close(x_re, st["x_re"], rtol=1e-3, atol=5e-4)
    assert set(grads) == set(st["grads"])
    for k in grads:
        ref = st["grads"][k]
        if k.endswith("blocks.0.0.bias"):
            wscale = float(st["grads"][k.replace(".bias", ".weight")].abs().max())
            assert float(grads[k].abs().max()) <= 1e-3 * wscale, k
            continue
        torch.testing.assert_close(grads[k], ref, rtol=2e-3, atol=1e-4 * float(ref.abs().max()) + 1e-9, msg=k)
    for k, v in st["buffers_after"].items():
        torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-6, msg=k)


# ---------------------------------------------------------------------------------------------------------------
# FC-latent variant (models/mymodel.py + utils/trainer_fc.py), SURVEY 8f NEXT-1
# ---------------------------------------------------------------------------------------------------------------
def fc_inputs(g):
    """Regenerate the fixture's inputs from its recorded CPU seed (the volumes are too large to commit)."""
    gen = torch.Generator().manual_seed(g["data_seed"])
    real = torch.rand(g["batch"], 1, 80, 96, 80, generator=gen)
    noise = torch.randn(g["batch"], g["z_ch"], generator=gen)
    assert float(real.double().sum()) == g["real_sum"] and torch.equal(noise, g["noise"])
    return real, noise


def bias_in_front_of_bn(k, tensors):
    """Conv3d biases followed by a train-mode BatchNorm3d: mathematically zero gradient, round-off noise only."""
    return k.endswith(".bias") and "last_block" not in k and tensors[k.replace(".bias", ".weight")].dim() == 5


def test_fc_eval_forward(golden_dir):
    g = _load(golden_dir, "fc_small.pt")
    real, _ = fc_inputs(g)
    cfg = O.FcCfg(*g["chans"], g["z_ch"])
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    mu, lv = O.encode(sd, real, cfg, False)
    x_re = O.decode(sd, mu, cfg, False)
    torch.testing.assert_close(mu, g["eval"]["mu"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lv, g["eval"]["logvar"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(x_re[:, :, ::4, ::4, ::4], g["eval"]["x_re_of_mu"]["sub"], rtol=1e-5, atol=1e-6)
    assert float(x_re.double().sum()) == pytest.approx(g["eval"]["x_re_of_mu"]["sum"], rel=1e-6)


def test_fc_soft_intro_step(golden_dir):
    g = _load(golden_dir, "fc_small.pt")
    st = g["step"]
    real, noise = fc_inputs(g)
    cfg = O.FcCfg(*g["chans"], g["z_ch"])
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    assert list(sd.keys()) == g["keys"]
    hp = O.StepHyper(**st["hyper"])
    terms, gE, gD = O.soft_intro_step_grads(sd, cfg, real, noise, st["eps"], None, hp)
    for k, v in st["terms"].items():
        assert terms[k] == pytest.approx(v, rel=5e-5, abs=1e-30), k
    assert set(gE) == set(st["gradsE"]) and set(gD) == set(st["gradsD"])
    assert not any(".block8." in k for k in gE)            # block8 is built but never run (mymodel.py:108-117)
    allref = {**st["gradsE"], **st["gradsD"]}
    for k, got in {**gE, **gD}.items():
        ref = allref[k]
        if bias_in_front_of_bn(k, allref):
            wscale = float(allref[k.replace(".bias", ".weight")].abs().max())
            assert float(got.abs().max()) <= 1e-3 * wscale and float(ref.abs().max()) <= 1e-3 * wscale, k
            continue
        torch.testing.assert_close(got, ref, rtol=2e-3, atol=1e-4 * float(ref.abs().max()) + 1e-12, msg=k)
    for k, v in st["buffers_after"].items():
        torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-6, msg=k)
    assert int(sd["encoder.block1.1.num_batches_tracked"]) == 5
    assert int(sd["decoder.block1.1.num_batches_tracked"]) == 8
    assert int(sd["encoder.block8.1.num_batches_tracked"]) == 0
