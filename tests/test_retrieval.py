"""Latent extraction + top-k similarity (SURVEY section 8f NEXT-4).
CPU: host logic against the kernel specification; GPU: the CUDA kernels through the C ABI against torch fp64."""
import pytest
import torch

import sivae_b200
from oracle import kernel_spec as S
from oracle import sivae_oracle as O
from tests.emu import emulated_kernels


def test_extract_latents_and_topk_cpu_emulated():
    torch.manual_seed(0)
    bs = [[4, 1, 2], [8, 1, 2], [8, 2, 2]]
    with emulated_kernels():
        net = sivae_b200.SoftIntroVAE(4, bs)
        net.apply(sivae_b200.init_weights_he)
        net.train()
        x = torch.rand(5, 1, 8, 8, 16)
        lat = sivae_b200.extract_latents(net, x, mode="mu", batch_size=2)
        assert net.training                                            # mode restored
        assert lat.shape == (5, 2) and lat.dtype == torch.float32      # latent 1x1x2 per volume
        # eval-mode encoder == the oracle's eval forward
        cfg = O.NetCfg.soft_intro(4, bs)
        mu, _ = O.encode({k: v.detach() for k, v in net.state_dict().items()}, x, cfg, False)
        assert torch.allclose(lat, mu.reshape(5, -1), rtol=1e-4, atol=1e-5)
        # loader-style input ((batch, label) pairs) and the reference's sampled latent
        z = sivae_b200.extract_latents(net, [(x[:3], None), (x[3:], None)], mode="z")
        assert z.shape == lat.shape and not torch.equal(z, lat)
        db = torch.randn(40, 12)
        sc, ix = sivae_b200.topk_similar(db[:7], db, k=3)
        assert torch.equal(ix[:, 0], torch.arange(7, dtype=torch.int32))   # every vector is its own best match
        assert torch.allclose(sc[:, 0], torch.ones(7), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["cosine", "l2"])
@pytest.mark.parametrize("nq,nd,dim,k", [(7, 50, 12, 5), (130, 1000, 1200, 10), (64, 64, 33, 32), (3, 70, 3, 1)])
def test_similarity_topk_gpu(metric, nq, nd, dim, k):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.manual_seed(1)
    q = torch.randn(nq, dim, device="cuda")
    d = torch.randn(nd, dim, device="cuda")
    d[: min(nq, nd) // 2] = q[: min(nq, nd) // 2]                       # exact matches must come first
    sc, ix = sivae_b200.topk_similar(q, d, k=k, metric=metric)
    rs, ri = S.similarity_topk(q, d, k, metric)
    assert torch.allclose(sc, rs, rtol=1e-4, atol=1e-4 * float(rs.abs().max()))
    # indices: identical except where two fp32 scores tie to within round-off
    same = ix == ri
    assert (~same).float().mean() < 0.01, float((~same).float().mean())
    for i in range(min(nq, nd) // 2):
        assert int(ix[i, 0]) == i


@pytest.mark.gpu
def test_extract_latents_gpu_matches_oracle_eval():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.manual_seed(2)
    bs = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
    net = sivae_b200.SoftIntroVAE(64, bs)
    net.apply(sivae_b200.init_weights_he)
    net.to("cuda")
    x = torch.rand(4, 1, 16, 24, 16, device="cuda")
    lat = sivae_b200.extract_latents(net, x, mode="mu", batch_size=2)
    cfg = O.NetCfg.soft_intro(64, bs)
    mu, _ = O.encode({k: v.detach() for k, v in net.state_dict().items()}, x, cfg, False)
    ref = mu.reshape(4, -1)
    assert lat.shape == ref.shape == (4, 12)
    err = float((lat - ref).abs().max())
    assert err <= 0.03 * float(ref.abs().max()) + 1e-3, (err, float(ref.abs().max()))
    sc, ix = sivae_b200.topk_similar(lat, lat, k=2)
    assert torch.equal(ix[:, 0].cpu(), torch.arange(4, dtype=torch.int32))
