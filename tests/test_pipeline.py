"""On-GPU input pipeline (SURVEY section 8f NEXT-3): BrainDataset._preprocess and the RandomAffine resampler.
CPU: a numpy restatement of utils/data_load.py:25-30 pins the specification; host logic runs on the emulated kernels.
GPU: the CUDA kernels through the C ABI against the specification (F.grid_sample for the resampler)."""
import numpy as np
import pytest
import torch

import sivae_b200
from sivae_b200 import kernels as K
from sivae_b200 import pipeline as P
from oracle import kernel_spec as S
from tests.emu import emulated_kernels


def _reference_preprocess(voxel: np.ndarray) -> np.ndarray:
    """utils/data_load.py:25-30, restated line by line (numpy, as the reference)."""
    cut_range = 4
    voxel = np.clip(voxel, 0, cut_range * np.std(voxel))
    voxel = (voxel - np.min(voxel)) / (np.max(voxel) - np.min(voxel))
    return voxel[np.newaxis, ].astype("f")


def _raw(b, d, h, w, dev="cpu"):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(b, d, h, w, generator=g) * 300 + 200                 # MRI-like: negative values and a long tail
    x[:, :2] = -50.0
    x[0, 3, 3, 3] = 1e5                                                  # an outlier far beyond 4 sigma
    return x.to(dev)


def test_preprocess_spec_matches_reference_numpy():
    x = _raw(3, 6, 7, 8)
    y, stats = S.preprocess_clip_minmax(x)
    for b in range(3):
        ref = _reference_preprocess(x[b].numpy())
        assert np.allclose(y[b].numpy(), ref[0], rtol=1e-5, atol=1e-6)
    assert float(y.min()) == 0.0 and float(y.max()) == 1.0


def test_pipeline_host_logic_cpu_emulated():
    with emulated_kernels():
        x = _raw(4, 8, 8, 8)[:, None]
        y = P.preprocess(x)
        assert y.shape == x.shape and float(y.min()) == 0.0 and float(y.max()) == 1.0
        g = torch.Generator().manual_seed(0)
        mats, applied = P.affine_matrices(64, (8, 8, 8), p=0.35, generator=g)
        assert 0.15 < float(applied.float().mean()) < 0.6
        ident = torch.tensor([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=torch.float32)
        assert all(torch.equal(mats[b], ident) for b in range(64) if not applied[b])
        m = mats[applied][0].reshape(3, 4)
        c = torch.tensor([3.5, 3.5, 3.5])
        assert torch.allclose(m[:, :3] @ c + m[:, 3], c, atol=1e-5)      # the centre is a fixed point
        det = float(torch.linalg.det(m[:, :3].double()))
        assert 1 / 1.1 ** 3 - 1e-6 <= det <= 1 / 0.9 ** 3 + 1e-6         # inverse of a scaling within 1 +- 0.1
        pipe = sivae_b200.GpuInputPipeline("cpu", train=True, p=1.0, seed=1)
        out = pipe(_raw(2, 8, 8, 8))
        assert out.shape == (2, 1, 8, 8, 8) and 0.0 <= float(out.min()) and float(out.max()) <= 1.0
        # identity matrices leave a volume unchanged
        x4 = _raw(2, 5, 6, 7)
        assert torch.allclose(S.affine_resample(x4, ident.repeat(2, 1)), x4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 6, 7, 8), (3, 16, 24, 16), (8, 80, 96, 80)])
def test_preprocess_gpu(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    x = _raw(*shape, dev="cuda")
    y, stats = K.preprocess_clip_minmax(x)
    ys, ss = S.preprocess_clip_minmax(x)
    assert torch.allclose(stats, ss, rtol=1e-5, atol=1e-3)
    assert torch.allclose(y, ys, rtol=1e-5, atol=1e-6)
    if shape[1] <= 16:
        for b in range(shape[0]):
            assert np.allclose(y[b].cpu().numpy(), _reference_preprocess(x[b].cpu().numpy())[0], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 6, 7, 8), (4, 16, 24, 16), (2, 40, 48, 40)])
def test_affine_resample_gpu(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    x = _raw(*shape, dev="cuda").clamp(-100, 2000)
    g = torch.Generator().manual_seed(3)
    mats, applied = P.affine_matrices(shape[0], shape[1:], p=0.75, generator=g)
    mats = mats.cuda()
    stats = K.volume_stats(x)
    y = K.affine_resample(x, mats, None, stats)
    ref = S.affine_resample(x, mats, None, stats)
    scale = float(x.abs().max())
    assert float((y - ref).abs().max()) <= 2e-3 * scale       # grid_sample's fp32 normalised coordinates differ in the last bits
    for b in range(shape[0]):
        if not applied[b]:
            assert torch.equal(y[b], x[b])                    # identity rows are exact
    pad = torch.full((shape[0],), 7.0, device="cuda")
    y2 = K.affine_resample(x, mats, pad, None)
    assert float((y2 - S.affine_resample(x, mats, pad, None)).abs().max()) <= 2e-3 * scale
