"""The UNMODIFIED reference training loop ``utils.my_trainer.train_soft_intro_vae`` (utils/my_trainer.py:147-508),
imported from the read-only checkout through oracle/ref_import.py, driving the drop-in model -- the integration
INTEGRATION.md section 2 describes: pass a ``sivae_b200.SoftIntroVAE`` and rebind the module-level loss functions.
libsivae.so is replaced by its executable specification (CPU).  The same loop is then run on the reference's own
``models.SoftIntroVAE``: identical seeds, init and data must give the same loss lists, the same artefacts (SURVEY
Q10 / Q17) and checkpoints that load in both directions.

Needs /root/reference (authoring container only): marked ``reference`` and skipped elsewhere."""
import os

import pytest
import torch

import sivae_b200
from sivae_b200 import trainer as T
from oracle import ref_import as R
from tests.emu import emulated_kernels

pytestmark = pytest.mark.reference
BS = [[4, 1, 2], [8, 1, 2], [8, 2, 2]]


def _loaders():
    g = torch.Generator().manual_seed(11)
    train = [(torch.rand(1, 1, 80, 96, 80, generator=g), torch.zeros(1))]      # the loop hard-codes 80x96x80 (Q7)
    val = [(torch.rand(1, 1, 80, 96, 80, generator=g), torch.zeros(1))]
    return train, val


def _no_dropout(net):
    for m in net.modules():        # the kernel specification takes explicit keep-masks only: both arms run without dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return net


def _run(ref_trainer, net, path):
    os.makedirs(os.path.join(path, "prams"))                                   # the reference never creates it (Q17)
    train, val = _loaders()
    return ref_trainer.train_soft_intro_vae(net, train, val, 1, device=torch.device("cpu"), path=path)


def test_unmodified_reference_loop_on_drop_in_model(tmp_path, monkeypatch):
    if not R.reference_available():
        pytest.skip("reference checkout not present")
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    ref_models, _, _, ref_trainer = R.import_reference()
    # arm 1: the reference's own model and loss functions
    p_ref = str(tmp_path / "ref") + "/"
    os.makedirs(p_ref)
    torch.manual_seed(123)          # constructor-time initialisation (conv biases) is not redone by the loop's init_weights_he
    out_ref = _run(ref_trainer, _no_dropout(ref_models.SoftIntroVAE(4, BS)), p_ref)
    # arm 2: the drop-in model; calc_kl / calc_reconstruction_loss are looked up as module globals at call time
    monkeypatch.setattr(ref_trainer, "calc_kl", T.calc_kl)
    monkeypatch.setattr(ref_trainer, "calc_reconstruction_loss", T.calc_reconstruction_loss)
    p_new = str(tmp_path / "new") + "/"
    os.makedirs(p_new)
    torch.manual_seed(123)          # same parameter-creation order -> the same constructor-time initialisation
    net = _no_dropout(sivae_b200.SoftIntroVAE(4, BS))
    with emulated_kernels():
        out_new = _run(ref_trainer, net, p_new)
    # same numbers: four lists, each epoch's value twice (Q10)
    assert len(out_ref) == len(out_new) == 4
    for a, b in zip(out_ref, out_new):
        assert len(a) == len(b) == 2 and a[0] == a[1] and b[0] == b[1]
        assert b[0] == pytest.approx(a[0], rel=2e-3), (out_ref, out_new)
    # same artefacts
    for p in (p_ref, p_new):
        assert open(p + "train_result.csv").read().strip() == "epoch,train_lossE,train_lossD,val_lossE,val_lossD"
        for f in ("loss.txt", "kl_losses.txt", "train_losses.txt", "val_losses.txt", "prams/S-IntroVAE_3898_epoch0.pth"):
            assert os.path.isfile(p + f), (p, f)
    sd_ref = torch.load(p_ref + "prams/S-IntroVAE_3898_epoch0.pth")
    sd_new = torch.load(p_new + "prams/S-IntroVAE_3898_epoch0.pth")
    assert list(sd_ref.keys()) == list(sd_new.keys())
    # checkpoints load in both directions, strictly
    ref_models.SoftIntroVAE(4, BS).load_state_dict(sd_new, strict=True)
    sivae_b200.SoftIntroVAE(4, BS).load_state_dict(sd_ref, strict=True)
    # 5 encoder / 8 decoder train-mode passes per iteration (Q15); eval passes do not count
    for sd in (sd_ref, sd_new):
        assert int(sd["encoder.blocks.0.1.num_batches_tracked"]) == 5
        assert int(sd["decoder.blocks.0.1.num_batches_tracked"]) == 8
    # after one E+D update the weights agree (Adam moves every weight by +-lr: agreement means the gradient signs agree)
    for k in ("encoder.blocks.1.0.block.0.weight", "decoder.blocks.3.0.block.4.weight", "encoder.mu.weight"):
        torch.testing.assert_close(sd_new[k], sd_ref[k], rtol=0, atol=4.1e-4, msg=k)
        assert float(((sd_new[k] - sd_ref[k]).abs() > 1e-5).float().mean()) < 0.05, k
    # the loop leaves the encoder frozen and the model on the CPU (Q12)
    assert not any(p.requires_grad for p in net.encoder.parameters())
    assert all(p.requires_grad for p in net.decoder.parameters())
