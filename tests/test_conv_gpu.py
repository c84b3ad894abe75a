"""tcgen05 convolution parity (GPU): fprop / dgrad / wgrad through the C ABI vs torch fp32 conv on the
same bf16-rounded inputs.  Tolerance: output is bf16-rounded (2^-8 relative) from fp32 accumulation."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import sivae_b200  # noqa: E402
from sivae_b200 import kernels as K  # noqa: E402
from oracle import kernel_spec as S  # noqa: E402

DEV = "cuda"
SHAPES = [
    (1, 8, 8, 16, 64, 64),        # exact 128-row tiles
    (2, 8, 12, 16, 64, 64),       # batch > 1
    (1, 10, 12, 10, 256, 256),    # headline latent resolution: 120-row tiles, 4 K-blocks per tap
    (1, 6, 8, 20, 256, 128),      # ragged W
    (1, 20, 24, 20, 64, 128),     # headline mid resolution
    (1, 5, 7, 9, 128, 64),        # every extent odd: overhanging boxes on all sides
    (2, 8, 16, 24, 128, 64),      # kd-fused kernel with two K blocks per tap (E2a dgrad shape class)
    (1, 6, 16, 8, 256, 64),       # ... four K blocks
    (1, 1, 1, 1, 64, 64),         # single voxel
    (1, 7, 16, 40, 64, 64),       # 8x16 patches of the persistent kw-slab kernel, odd depth (half-empty depth pair)
    (2, 4, 24, 32, 64, 128),      # ... two output-channel blocks, batch 2
    (1, 6, 30, 44, 64, 64),       # ... ragged W and H (clipped stores, OOB-filled loads)
    (3, 12, 32, 24, 64, 64),      # ... more work items than one wave of a small grid would take
]
# Cin = 64 shapes run several ways: the library's own choice ("auto": the persistent kd-fused kernel for well-tiling
# 64 -> 64 shapes, else tap-by-tap), the kd-fused kernel forced for every 64 -> 64 shape incl. ragged / tiny ones
# (SIVAE_CONV_KD=force), tap-by-tap forced (SIVAE_CONV_KD=0) and the opt-in kw-slab kernel forced (SIVAE_CONV_KW=force).
KW_MODES = ["auto", "force", "0"]
CONV_MODES = {"auto": {}, "kd_force": {"SIVAE_CONV_KD": "force"}, "tapwise": {"SIVAE_CONV_KD": "0"},
              "kw_force": {"SIVAE_CONV_KD": "0", "SIVAE_CONV_KW": "force"},
              # the 6-stage operand ring of the Cout % 128 == 0 kernel (taken by itself only for grids <= one CTA per SM)
              "deep_ring": {"SIVAE_CONV_KD": "0", "SIVAE_DEEP_RING": "1"},
              "shallow_ring": {"SIVAE_CONV_KD": "0", "SIVAE_DEEP_RING": "0"},
              # N = 256 tiles with the 4-stage ring (the library's choice for Cout % 256 == 0 grids of more than one wave
              # of N = 128 CTAs) forced for every Cout % 256 == 0 shape, and the legacy 3-stage form
              "n256_ring4": {"SIVAE_CONV_KD": "0", "SIVAE_N256": "4"},
              "n256_ring3": {"SIVAE_CONV_KD": "0", "SIVAE_N256": "1"}}



@pytest.fixture(params=list(CONV_MODES))
def kwmode(request):
    import os
    keys = ("SIVAE_CONV_KD", "SIVAE_CONV_KW", "SIVAE_DEEP_RING", "SIVAE_N256")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(CONV_MODES[request.param])
    yield request.param
    for k in keys:
        os.environ.pop(k, None)
        if old[k] is not None:
            os.environ[k] = old[k]


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(7)
    yield
    torch.cuda.synchronize()


def _mk(n, d, h, w, ci, co):
    x = torch.randn(n, d, h, w, ci, device=DEV).to(torch.bfloat16)
    wt = torch.randn(co, ci, 3, 3, 3, device=DEV) * (2.0 / (27 * ci)) ** 0.5
    return x, wt


def _check_bf16(got, ref, what):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all(), what
    err = (got - ref).abs()
    tol = 2 ** -7 * ref.abs() + 2 ** -7 * float(ref.abs().mean()) + 1e-6
    assert not (err > tol).any(), f"{what}: {int((err > tol).sum())}/{err.numel()} off, max {float(err.max()):.4e}"


def test_pack_weights_exact():
    wt = torch.randn(128, 64, 3, 3, 3, device=DEV)
    wf, wd = K.pack_conv3_weights(wt)
    wf_s, wd_s = S.pack_conv3_weights(wt)
    assert torch.equal(wf, wf_s) and torch.equal(wd, wd_s)


@pytest.mark.parametrize("shape", SHAPES)
def test_fprop(shape, kwmode):
    # kd-fused kernel: any Cin with Cout = 64; kw-slab kernel: Cin = 64
    if (kwmode == "kw_force" and shape[4] != 64) or (kwmode in ("kd_force", "tapwise") and shape[5] != 64 and shape[4] != 64):
        pytest.skip("no kernel choice for this channel combination")
    x, wt = _mk(*shape)
    wf, _ = K.pack_conv3_weights(wt)
    _check_bf16(K.conv3_igemm(x, wf), S.conv3_igemm(x, wf), f"fprop {shape}")


@pytest.mark.parametrize("shape", SHAPES)
def test_dgrad_is_conv_transpose(shape, kwmode):
    """dgrad = the same kernel on the flipped/transposed pack; checked against autograd of F.conv3d."""
    n, d, h, w, ci, co = shape
    # the kernel sees GEMM-K = co, GEMM-N = ci
    if (kwmode == "kw_force" and co != 64) or (kwmode in ("kd_force", "tapwise") and ci != 64 and co != 64):
        pytest.skip("no kernel choice for this channel combination")
    x, wt = _mk(*shape)
    dy = torch.randn(n, d, h, w, co, device=DEV).to(torch.bfloat16)
    _, wd = K.pack_conv3_weights(wt)
    got = K.conv3_igemm(dy, wd)
    wq = wt.to(torch.bfloat16).float()
    xin = torch.zeros(n, ci, d, h, w, device=DEV, requires_grad=True)
    torch.nn.functional.conv3d(xin, wq, None, 1, 1).backward(dy.float().permute(0, 4, 1, 2, 3))
    _check_bf16(got, xin.grad.permute(0, 2, 3, 4, 1), f"dgrad {shape}")


@pytest.fixture(params=KW_MODES)
def wgmode(request):
    import os
    old = os.environ.get("SIVAE_WGRAD_KW")
    if request.param == "auto":
        os.environ.pop("SIVAE_WGRAD_KW", None)
    else:
        os.environ["SIVAE_WGRAD_KW"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("SIVAE_WGRAD_KW", None)
    else:
        os.environ["SIVAE_WGRAD_KW"] = old


@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad(shape, wgmode):
    n, d, h, w, ci, co = shape
    if wgmode != "auto" and ci != 64:
        pytest.skip("kernel choice only exists for Cin = 64")
    x, _ = _mk(*shape)
    dy = torch.randn(n, d, h, w, co, device=DEV).to(torch.bfloat16)
    got = K.conv3_wgrad(x, dy)
    ref = S.conv3_wgrad(x, dy)
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 2e-3 * scale + 1e-5, f"wgrad {shape}: max err {err:.4e} vs scale {scale:.4e}"


def test_linearity_at_headline_size():
    """Size-independent property at the full 80x96x80 resolution: conv(a*x1 + x2) = a*conv(x1) + conv(x2) and
    agreement with cuDNN fp32 on a random sub-block (the full fp32 reference would be slow, not wrong)."""
    n, d, h, w, ci, co = 1, 80, 96, 80, 64, 64
    x1, wt = _mk(n, d, h, w, ci, co)
    wf, _ = K.pack_conv3_weights(wt)
    y1 = K.conv3_igemm(x1, wf)
    y2 = K.conv3_igemm((x1.float() * 2).to(torch.bfloat16), wf)       # exact scaling by 2 in bf16
    assert torch.equal(y2.float(), y1.float() * 2)
    sub = S.conv3_igemm(x1[:, 30:50].contiguous(), wf)                # interior planes 31..48 are halo-free
    _check_bf16(y1[:, 31:49], sub[:, 1:19], "headline sub-block")


# ---------------------------------------------------------------- Upsample(2) folded into the convolution
UP_SHAPES = [
    (1, 4, 4, 8, 64, 64),       # low-res extents; exact tiles
    (2, 5, 6, 5, 128, 64),      # D3b-like channel change, ragged
    (1, 10, 12, 10, 256, 128),  # D2b at the headline resolution
    (1, 3, 5, 7, 64, 64),       # all odd
    (1, 1, 1, 1, 64, 128),      # single low-res voxel -> 2x2x2 outputs
]


def _upsample(x):  # NDHWC nearest x2
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)


def test_pack_upconv_weights():
    wt = torch.randn(128, 64, 3, 3, 3, device=DEV)
    wup, wupT = K.pack_upconv3_weights(wt)
    sup, supT = S.pack_upconv3_weights(wt)
    # fp32 summation order may differ by an ulp before the bf16 rounding
    assert float((wup.float() - sup.float()).abs().max()) <= 2 ** -7 * float(sup.float().abs().max())
    assert (wup != sup).float().mean() < 1e-3
    assert torch.equal(wupT, wup.transpose(1, 2).contiguous())


@pytest.fixture(params=["auto", "force", "0"])
def upmode(request):
    """persistent shared-tile kernel: library's choice / forced for every Cout = 64 shape / tap-by-tap forced"""
    import os
    old = os.environ.pop("SIVAE_UPCONV_FUSED", None)
    if request.param != "auto":
        os.environ["SIVAE_UPCONV_FUSED"] = request.param
    yield request.param
    os.environ.pop("SIVAE_UPCONV_FUSED", None)
    if old is not None:
        os.environ["SIVAE_UPCONV_FUSED"] = old


@pytest.mark.parametrize("shape", UP_SHAPES + [(2, 6, 16, 24, 64, 64), (1, 3, 32, 8, 128, 64), (3, 5, 20, 13, 64, 64)])
def test_upconv_fprop_bn_fused_stats(shape, upmode):
    """sivae_upconv3_fprop_bn = sivae_upconv3_fprop + sivae_bn_train_coeffs (statistics from the epilogue of the
    persistent kernel when it runs; ragged low-res tiles masked)."""
    n, d, h, w, ci, co = shape
    x, wt = _mk(*shape)
    wup, _ = K.pack_upconv3_weights(wt)
    gamma, beta = torch.rand(co, device=DEV) + 0.5, torch.randn(co, device=DEV)
    rm1, rv1, n1 = torch.zeros(co, device=DEV), torch.ones(co, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    rm2, rv2, n2 = rm1.clone(), rv1.clone(), n1.clone()
    y, *coef = K.upconv3_fprop_bn(x, wup, gamma, beta, rm1, rv1, n1, 0.1, 1e-5)
    y_ref = K.upconv3_fprop(x, wup)
    ref = K.bn_train_coeffs(y_ref, gamma, beta, rm2, rv2, n2, 0.1, 1e-5)
    assert torch.equal(y, y_ref)
    _check_bf16(y, S.upconv3_fprop(x, wup), f"upconv fprop (bn entry) vs folded spec {shape}")
    for a, b, name in zip(coef, ref, ("mean", "invstd", "scale", "shift")):
        assert torch.allclose(a, b, rtol=3e-5, atol=3e-6), (name, float((a - b).abs().max()))
    assert torch.allclose(rm1, rm2, rtol=3e-5, atol=3e-6) and torch.allclose(rv1, rv2, rtol=3e-5, atol=3e-6)


@pytest.mark.parametrize("shape", UP_SHAPES)
def test_upconv_fprop(shape, upmode):
    n, d, h, w, ci, co = shape
    x, wt = _mk(*shape)
    wup, _ = K.pack_upconv3_weights(wt)
    got = K.upconv3_fprop(x, wup)
    _check_bf16(got, S.upconv3_fprop(x, wup), f"upconv fprop vs folded spec {shape}")
    # and against the definition: conv3d over the materialised nearest-upsampled tensor with the fp32 weights
    # (differs only by the bf16 rounding of the pre-summed weights)
    ref = torch.nn.functional.conv3d(_upsample(x).float().permute(0, 4, 1, 2, 3), wt, None, 1, 1).permute(0, 2, 3, 4, 1)
    err = float((got.float() - ref).abs().max())
    assert err <= 0.03 * float(ref.abs().max()) + 1e-3, (shape, err)


@pytest.mark.parametrize("shape", UP_SHAPES)
def test_upconv_dgrad(shape):
    n, d, h, w, ci, co = shape
    _, wt = _mk(*shape)
    dy = torch.randn(n, 2 * d, 2 * h, 2 * w, co, device=DEV).to(torch.bfloat16)
    _, wupT = K.pack_upconv3_weights(wt)
    _check_bf16(K.upconv3_dgrad(dy, wupT), S.upconv3_dgrad(dy, wupT), f"upconv dgrad {shape}")


@pytest.mark.parametrize("mode", ["auto", "force", "0"])
@pytest.mark.parametrize("shape", UP_SHAPES + [(2, 6, 16, 24, 64, 64), (1, 5, 20, 13, 64, 128), (3, 4, 32, 8, 64, 64)])
def test_upconv_wgrad(shape, mode):
    """mode: persistent tall-box kernel by the library's choice / forced for every Cin = 64 shape (ragged patches, odd
    depth) / generic kernel forced (SIVAE_UPWGRAD_TALL)."""
    import os
    n, d, h, w, ci, co = shape
    if mode != "auto" and ci != 64:
        pytest.skip("kernel choice only exists for Cin = 64")
    old = os.environ.pop("SIVAE_UPWGRAD_TALL", None)
    if mode != "auto":
        os.environ["SIVAE_UPWGRAD_TALL"] = mode
    try:
        _upconv_wgrad_check(shape)
    finally:
        os.environ.pop("SIVAE_UPWGRAD_TALL", None)
        if old is not None:
            os.environ["SIVAE_UPWGRAD_TALL"] = old


def _upconv_wgrad_check(shape):
    n, d, h, w, ci, co = shape
    x, _ = _mk(*shape)
    dy = torch.randn(n, 2 * d, 2 * h, 2 * w, co, device=DEV).to(torch.bfloat16)
    got = K.upconv3_wgrad(x, dy)
    ref = S.upconv3_wgrad(x, dy)
    assert torch.isfinite(got).all()
    err, scale = float((got - ref).abs().max()), float(ref.abs().max())
    assert err <= 2e-3 * scale + 1e-5, f"upconv wgrad {shape}: max err {err:.4e} vs scale {scale:.4e}"


# ---------------------------------------------------------------- convolution + fused BatchNorm statistics
@pytest.mark.parametrize("shape", [(1, 8, 16, 16, 64, 64), (2, 6, 30, 44, 64, 64), (1, 7, 16, 40, 64, 64),
                                   (3, 12, 32, 24, 64, 64), (1, 5, 7, 9, 128, 64), (1, 10, 12, 10, 256, 256)])
@pytest.mark.parametrize("mode", ["auto", "kd_force"])
def test_conv_bn_fused_stats(shape, mode):
    """sivae_conv3_igemm_bn = sivae_conv3_igemm + sivae_bn_train_coeffs: with the persistent kd-fused kernel
    (forced here also for ragged shapes: partial tiles must be masked out of the sums) the statistics come from the
    convolution epilogue; they must equal the separate pass over the stored tensor."""
    import os
    old = os.environ.pop("SIVAE_CONV_KD", None)
    if mode == "kd_force":
        os.environ["SIVAE_CONV_KD"] = "force"
    try:
        n, d, h, w, ci, co = shape
        x, wt = _mk(*shape)
        wf, _ = K.pack_conv3_weights(wt)
        gamma, beta = torch.rand(co, device=DEV) + 0.5, torch.randn(co, device=DEV)
        rm1, rv1, nbt1 = torch.zeros(co, device=DEV), torch.ones(co, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
        rm2, rv2, nbt2 = rm1.clone(), rv1.clone(), nbt1.clone()
        y, mean, invstd, scale, shift = K.conv3_igemm_bn(x, wf, gamma, beta, rm1, rv1, nbt1, 0.1, 1e-5)
        y_ref = K.conv3_igemm(x, wf)
        ref = K.bn_train_coeffs(y_ref, gamma, beta, rm2, rv2, nbt2, 0.1, 1e-5)
        assert torch.equal(y, y_ref)
        for a, b, name in zip((mean, invstd, scale, shift), ref, ("mean", "invstd", "scale", "shift")):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (name, float((a - b).abs().max()))
        assert torch.allclose(rm1, rm2, rtol=2e-5, atol=2e-6) and torch.allclose(rv1, rv2, rtol=2e-5, atol=2e-6)
        assert int(nbt1) == 1 and int(nbt2) == 1
        # and against the definition
        yf = y.float().reshape(-1, co)
        assert torch.allclose(mean, yf.mean(0), rtol=1e-4, atol=1e-5)
        assert torch.allclose(invstd, (yf.var(0, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    finally:
        os.environ.pop("SIVAE_CONV_KD", None)
        if old is not None:
            os.environ["SIVAE_CONV_KD"] = old
