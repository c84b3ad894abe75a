"""Per-kernel parity (GPU): every function of the C ABI, called through ``sivae_b200.kernels``, against
its executable specification ``oracle/kernel_spec.py`` evaluated by torch on the same device in fp32
(TF32 off).  Tolerances: bf16 outputs -> 2^-8 relative rounding + accumulation slack; fp32 outputs ->
1e-5 relative (stated per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import sivae_b200  # noqa: E402
from sivae_b200 import kernels as K  # noqa: E402
from oracle import kernel_spec as S  # noqa: E402


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    K.device_check()
    torch.manual_seed(1234)
    yield
    torch.cuda.synchronize()


DEV = "cuda"


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(torch.bfloat16)


def assert_bf16_close(got, ref, what, rel=2 ** -7, slack=None):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    tol = rel * ref.abs() + (slack if slack is not None else rel * float(ref.abs().mean()) + 1e-6)
    bad = (got - ref).abs() > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {float((got - ref).abs().max()):.4e}"


def assert_f32_close(got, ref, what, rtol=1e-4):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = float(ref.abs().max()) + 1e-30
    err = float((got - ref).abs().max())
    assert err <= rtol * scale, f"{what}: max err {err:.4e} vs scale {scale:.4e}"


# ------------------------------------------------------------------------------------------- pointwise
@pytest.mark.parametrize("C", [64, 128, 256])
def test_bn_train_coeffs(C):
    y = bf(2, 6, 8, 10, C, scale=2.0) + 0.5
    gamma = torch.rand(C, device=DEV) + 0.5
    beta = torch.randn(C, device=DEV)
    rm, rv = torch.randn(C, device=DEV), torch.rand(C, device=DEV) + 0.5
    nbt = torch.tensor(3, device=DEV)
    rm2, rv2, nbt2 = rm.clone(), rv.clone(), nbt.clone()
    got = K.bn_train_coeffs(y, gamma, beta, rm, rv, nbt, 0.1, 1e-5)
    ref = S.bn_train_coeffs(y, gamma, beta, rm2, rv2, nbt2, 0.1, 1e-5)
    for g, r, n in zip(got, ref, ("mean", "invstd", "scale", "shift")):
        assert_f32_close(g, r, n, 2e-5)
    assert_f32_close(rm, rm2, "running_mean", 2e-5)
    assert_f32_close(rv, rv2, "running_var", 2e-5)
    assert int(nbt) == 4


@pytest.mark.parametrize("resample", [0, 1, 2])
@pytest.mark.parametrize("with_res,with_mask", [(False, False), (True, False), (False, True)])
def test_bn_act_fwd_bwd(resample, with_res, with_mask):
    C = 64
    y = bf(2, 4, 6, 8, C)
    res = bf(2, 4, 6, 8, C) if with_res else None
    mask = (torch.rand(2, 4, 6, 8, C, device=DEV) > 0.35).to(torch.uint8) if with_mask else None
    p = 0.35 if with_mask else 0.0
    gamma = torch.rand(C, device=DEV) + 0.5
    beta = torch.randn(C, device=DEV) * 0.3
    mean, invstd, scale, shift = S.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
    out = K.bn_act_fwd(y, scale, shift, res, 0.2, resample, mask, p, 0)
    ref = S.bn_act_fwd(y, scale, shift, res, 0.2, resample, mask, p, 0)
    assert_bf16_close(out, ref, "bn_act_fwd")
    g = bf(*ref.shape)
    got = K.bn_act_bwd(g, y, res, mean, invstd, gamma, beta, 0.2, resample, mask, p, 0, need_dres=with_res)
    exp = S.bn_act_bwd(g, y, res, mean, invstd, gamma, beta, 0.2, resample, mask, p, 0, need_dres=with_res)
    assert_bf16_close(got[0], exp[0], "dconv", slack=0.02 * float(exp[0].float().abs().mean()) + 1e-6)
    if with_res:
        assert_bf16_close(got[1], exp[1], "dres")
    assert_f32_close(got[2], exp[2], "dgamma", 1e-3)
    assert_f32_close(got[3], exp[3], "dbeta", 1e-3)


@pytest.mark.parametrize("shape,C", [((2, 4, 6, 8), 64), ((3, 17, 19, 23), 64), ((1, 9, 10, 11), 256), ((4, 40, 48, 40), 64)])
def test_bn_act_plain_fwd_equals_general_kernel(shape, C):
    """The streaming kernel taken when there is no residual / dropout / resampling (bn_act_plain_fwd_kernel, 8 loads in
    flight per thread) must return exactly what the general kernel returns: the general kernel is forced by a zero
    residual, which adds +0.0 to every value (only the sign of an exact zero can differ, and compares equal)."""
    y = bf(*shape, C, scale=1.5)
    scale, shift = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.4
    plain = K.bn_act_fwd(y, scale, shift, None, 0.2, 0)
    general = K.bn_act_fwd(y, scale, shift, torch.zeros_like(y), 0.2, 0)
    assert torch.equal(plain, general)
    assert_bf16_close(plain, S.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.0, 0), "bn_act_plain_fwd vs spec")


def test_bn_act_relu_slope_zero():
    C = 128
    y = bf(1, 4, 4, 4, C)
    scale, shift = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    assert_bf16_close(K.bn_act_fwd(y, scale, shift, None, 0.0, 0), S.bn_act_fwd(y, scale, shift, None, 0.0, 0), "relu")


def test_philox_dropout_is_consistent_and_calibrated():
    C, p = 64, 0.35
    y = torch.ones(2, 8, 8, 8, C, device=DEV, dtype=torch.bfloat16)
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    a = K.bn_act_fwd(y, one, zero, None, 0.2, 0, None, p, 1234567)
    b = K.bn_act_fwd(y, one, zero, None, 0.2, 0, None, p, 1234567)
    c = K.bn_act_fwd(y, one, zero, None, 0.2, 0, None, p, 7654321)
    assert torch.equal(a, b) and not torch.equal(a, c)
    keep = (a != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01, keep
    assert torch.allclose(a[a != 0].float(), torch.tensor(1 / (1 - p), device=DEV), rtol=1e-2)
    # backward regenerates the same mask: with identity BN and positive inputs dt = keep/(1-p), so
    # sum(dbeta)/numel = E[keep/(1-p)] ~ 1 and dconv vanishes exactly where the forward output is zero
    g = torch.ones_like(a)
    dconv, dres, dgamma, dbeta = K.bn_act_bwd(g, y, None, zero, one, one, zero, 0.2, 0, None, p, 1234567,
                                              need_dres=True)
    assert torch.equal(dres != 0, a != 0)
    assert abs(float(dbeta.sum()) / a.numel() - 1.0) < 0.02
    # the lean kernels (taken when no dres is requested) regenerate the same Philox mask as the generic ones
    yr = bf(*y.shape)
    gr = bf(*y.shape)
    gam, bet = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    mean, invstd, _, _ = S.bn_train_coeffs(yr, gam, bet, None, None, None, 0.1, 1e-5)
    d1, _, dg1, db1 = K.bn_act_bwd(gr, yr, None, mean, invstd, gam, bet, 0.2, 0, None, p, 99, need_dres=True)
    d2, _, dg2, db2 = K.bn_act_bwd(gr, yr, None, mean, invstd, gam, bet, 0.2, 0, None, p, 99, need_dres=False)
    assert_bf16_close(d2, d1, "lean vs generic dconv under Philox dropout")
    assert_f32_close(dg2, dg1, "dgamma", 1e-4)
    assert_f32_close(db2, db1, "dbeta", 1e-4)


@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 64), (1, 5, 7, 3, 128), (3, 2, 2, 2, 256)])
def test_dropout_keep_bit_store_equals_philox_regeneration(shape):
    """The keep-bit store (forward writes 1 bit per element, backward reads it) must reproduce bit for bit what the
    Philox-regenerating kernels compute for the same seed: same forward output, same dconv / dgamma / dbeta; the
    stored bits are exactly the non-zero pattern of the forward output, and the kernel specification fed with the
    decoded bits as an explicit mask agrees."""
    C, p, seed = shape[-1], 0.35, 0xC0FFEE
    y, g = bf(*shape), bf(*shape)
    gam, bet = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    mean, invstd, scale, shift = S.bn_train_coeffs(y, gam, bet, None, None, None, 0.1, 1e-5)
    bits = torch.zeros(y.numel() // 8, dtype=torch.uint8, device=DEV)
    a_ref = K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, p, seed)
    a = K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, p, seed, keep_bits=bits)
    assert torch.equal(a, a_ref)
    keep = S._unpack_keep_bits(bits, y.shape)
    t = y.float() * scale + shift
    assert torch.equal(keep | (t == 0), (a != 0) | (t == 0))
    assert abs(float(keep.float().mean()) - (1 - p)) < 0.05
    d_ref, _, dg_ref, db_ref = K.bn_act_bwd(g, y, None, mean, invstd, gam, bet, 0.2, 0, None, p, seed)
    d, _, dg, db = K.bn_act_bwd(g, y, None, mean, invstd, gam, bet, 0.2, 0, None, p, seed, keep_bits=bits)
    assert torch.equal(d, d_ref) and torch.equal(dg, dg_ref) and torch.equal(db, db_ref)
    exp = S.bn_act_bwd(g, y, None, mean, invstd, gam, bet, 0.2, 0, keep.to(torch.uint8), p, 0)
    assert_bf16_close(d, exp[0], "dconv vs spec", slack=0.02 * float(exp[0].float().abs().mean()) + 1e-6)
    assert_f32_close(dg, exp[2], "dgamma", 1e-3)
    assert_f32_close(db, exp[3], "dbeta", 1e-3)


# ------------------------------------------------------------------------------------------- thin convs
@pytest.mark.parametrize("C,T", [(64, 27), (64, 1), (128, 27), (256, 1)])
@pytest.mark.parametrize("flip", [False, True])
def test_c1_to_cn(C, T, flip):
    x1 = torch.randn(2, 5, 6, 9, device=DEV)
    w = torch.randn(C, T, device=DEV) * 0.3
    b = torch.randn(C, device=DEV)
    got = K.c1_to_cn(x1, w, b, flip)
    ref = S.c1_to_cn(x1, w, b, flip)
    assert_bf16_close(got, ref, "c1_to_cn")
    x2 = torch.randn(2, 5, 6, 9, device=DEV)
    acc = K.c1_to_cn(x2, w, None, flip, out=got.clone())
    ref2 = S.c1_to_cn(x2, w, None, flip, out=ref.clone())
    assert_bf16_close(acc, ref2, "c1_to_cn accumulate", rel=2 ** -6)


@pytest.mark.parametrize("shape", [(2, 5, 6, 9), (1, 16, 16, 32), (3, 7, 20, 40), (8, 20, 24, 16)])
@pytest.mark.parametrize("C,T", [(64, 27), (128, 27), (256, 1)])
def test_c1_to_cn_bn_fused_stats(shape, C, T):
    """Stem convolution + BatchNorm coefficients in one call: Conv3d(1,64,3) takes the channel sums in the convolution
    epilogue (ragged tiles masked), other widths run the two passes; both must match conv followed by the stats pass."""
    x1 = torch.randn(*shape, device=DEV)
    w = torch.randn(C, T, device=DEV) * 0.3
    b = torch.randn(C, device=DEV)
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    rm1, rv1, n1 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    rm2, rv2, n2 = rm1.clone(), rv1.clone(), n1.clone()
    y, *coef = K.c1_to_cn_bn(x1, w, b, gamma, beta, rm1, rv1, n1, 0.1, 1e-5)
    y_ref = K.c1_to_cn(x1, w, b)
    ref = K.bn_train_coeffs(y_ref, gamma, beta, rm2, rv2, n2, 0.1, 1e-5)
    assert torch.equal(y, y_ref)
    for a, r, name in zip(coef, ref, ("mean", "invstd", "scale", "shift")):
        assert_f32_close(a, r, name, 3e-5)
    assert_f32_close(rm1, rm2, "running_mean", 3e-5)
    assert_f32_close(rv1, rv2, "running_var", 3e-5)
    assert int(n1) == 1


@pytest.mark.parametrize("C,T", [(64, 27), (64, 1), (128, 27), (128, 1), (256, 1)])
@pytest.mark.parametrize("flip,act", [(False, 0), (True, 0), (False, 1)])
def test_cn_to_c1(C, T, flip, act):
    x = bf(2, 5, 7, 37, C)   # W=37: ragged 32-wide bricks
    w = torch.randn(C, T, device=DEV) * 0.1
    b = torch.randn(1, device=DEV)
    mask = (torch.rand(2, 5, 7, 37, device=DEV) > 0.35).to(torch.uint8) if act else None
    got = K.cn_to_c1(x, w, b, flip, act, mask, 0.35 if act else 0.0, 0)
    # the 3x3x3 case runs on tcgen05 with split-bf16 weights (hi + lo ~ 16 mantissa bits, fp32 accumulation);
    # the 1x1 case stays fp32 SIMT
    ref = S.cn_to_c1(x, w, b, flip, act, mask, 0.35 if act else 0.0, 0)
    assert_f32_close(got, ref, "cn_to_c1", 1e-4 if T == 27 else 2e-5)
    # the direct (SIMT) C ABI entry point stays available and exact in fp32
    import ctypes
    lib = K.load_library()
    y = torch.empty_like(ref)
    rc = lib.sivae_cn_to_c1(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 2, 5, 7, 37, C, T, int(flip), act,
                            mask.data_ptr() if mask is not None else None, 0.35 if act else 0.0, 0,
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    assert_f32_close(y, S.cn_to_c1(x, w, b, flip, act, mask, 0.35 if act else 0.0, 0), "cn_to_c1 (SIMT)", 2e-5)


@pytest.mark.parametrize("act", [0, 1])
def test_cn_to_c1_many_planes_per_cta(act, monkeypatch):
    """conv3_to1_halo_kernel over a geometry where every CTA walks several work items of 17 input planes each: the two
    converter and the two gather warp groups alternate planes across item boundaries and the 5-deep P ring wraps dozens
    of times.  Checked against the specification, against the tap-by-tap tcgen05 kernel (SIVAE_TO1_TAPWISE) including
    the Philox keep pattern, and for run-to-run identity."""
    x = bf(3, 44, 50, 70, 64)
    w = torch.randn(64, 27, device=DEV) * 0.1
    b = torch.randn(1, device=DEV)
    p, seed = (0.35, 4242) if act else (0.0, 0)
    got = K.cn_to_c1(x, w, b, False, act, None, p, seed)
    again = K.cn_to_c1(x, w, b, False, act, None, p, seed)
    assert torch.equal(got, again)
    monkeypatch.setenv("SIVAE_TO1_TAPWISE", "1")
    tap = K.cn_to_c1(x, w, b, False, act, None, p, seed)
    monkeypatch.delenv("SIVAE_TO1_TAPWISE")
    # same Philox keep pattern; only pre-activations within rounding of zero may differ in being clipped by the ReLU
    assert float(((got == 0) != (tap == 0)).float().mean()) < 1e-4
    assert_f32_close(got, tap, "cn_to_c1 halo vs tap-by-tap", 1e-4)
    if act == 0:
        assert_f32_close(got, S.cn_to_c1(x, w, b, False, 0, None, 0.0, 0), "cn_to_c1 vs spec", 1e-4)


@pytest.mark.parametrize("C,T", [(64, 27), (64, 1), (128, 27), (256, 1)])
@pytest.mark.parametrize("flip", [False, True])
def test_wgrad_c1(C, T, flip):
    xc = bf(2, 6, 5, 11, C)
    x1 = torch.randn(2, 6, 5, 11, device=DEV)
    got = K.wgrad_c1(xc, x1, T, flip)
    ref = S.wgrad_c1(xc, x1, T, flip)
    for g, r, n in zip(got, ref, ("dw", "sum_c", "sum_1")):
        assert_f32_close(g, r, n, 1e-4)


def test_relu_drop_bwd():
    g, out = torch.randn(1000, device=DEV), torch.relu(torch.randn(1000, device=DEV))
    assert torch.equal(K.relu_drop_bwd(g, out, 0.35), S.relu_drop_bwd(g, out, 0.35))


# ------------------------------------------------------------------------------------------- latent / loss
def test_reparam_bit_exact_and_backward():
    mu, lv, eps = (torch.randn(8, 1200, device=DEV) for _ in range(3))
    z = K.reparam_fwd(mu, lv, eps)
    assert torch.equal(z, mu + eps * torch.exp(0.5 * lv))            # bit-exact vs torch on the same device
    assert torch.equal(K.reparam_fwd(mu, lv, 0.1), mu + 0.1 * torch.exp(0.5 * lv))
    dz = torch.randn_like(mu)
    dmu, dlv = K.reparam_bwd(dz, lv, eps)
    rmu, rlv = S.reparam_bwd(dz, lv, eps)
    assert torch.equal(dmu, rmu)
    assert_f32_close(dlv, rlv, "dlogvar", 1e-6)


def test_reparam_sampler_statistics_and_determinism():
    """sivae_reparam_draw_fwd: eps ~ N(0,1) drawn in the kernel (Philox + Box-Muller).  Same seed -> same draw, other seed
    -> another; moments of 4M draws within sampling error; z obeys the bit-exact reparameterisation rule given the
    returned eps; odd lengths (tail of the 4-element Philox block) are covered."""
    n = 1 << 22
    mu = torch.randn(n, device=DEV)
    lv = torch.randn(n, device=DEV) * 0.5
    z1, e1 = K.reparam_draw_fwd(mu, lv, 12345)
    z2, e2 = K.reparam_draw_fwd(mu, lv, 12345)
    z3, e3 = K.reparam_draw_fwd(mu, lv, 54321)
    assert torch.equal(e1, e2) and torch.equal(z1, z2) and not torch.equal(e1, e3)
    assert torch.isfinite(e1).all()
    m, v = float(e1.mean()), float(e1.var())
    skew, kurt = float((e1 ** 3).mean()), float((e1 ** 4).mean())
    assert abs(m) < 3e-3 and abs(v - 1) < 5e-3 and abs(skew) < 1e-2 and abs(kurt - 3) < 3e-2, (m, v, skew, kurt)
    assert float(e1.abs().max()) > 4.5                              # tails are there (P(|x| > 4.5) * 4M = 28)
    assert abs(float((e1[0::4] * e1[1::4]).mean())) < 3e-3          # the two outputs of a Box-Muller pair are uncorrelated
    assert abs(float((e1[:-1] * e1[1:]).mean())) < 3e-3
    assert torch.equal(z1, K.reparam_fwd(mu, lv, e1))               # same three rounded operations as the injected path
    for odd in (1, 3, 5, 1201):
        zo, eo = K.reparam_draw_fwd(mu[:odd].contiguous(), lv[:odd].contiguous(), 777)
        assert zo.shape == (odd,) and torch.isfinite(eo).all()
        assert torch.equal(eo, K.reparam_draw_fwd(mu[:1201].contiguous(), lv[:1201].contiguous(), 777)[1][:odd])


@pytest.mark.parametrize("B,n", [(8, 1200), (3, 7), (1, 9600)])
def test_kl(B, n):
    mu, lv = torch.randn(B, n, device=DEV), torch.randn(B, n, device=DEV) * 0.5
    assert_f32_close(K.kl_persample_fwd(mu, lv), S.kl_persample_fwd(mu, lv), "kl", 1e-5)
    g = torch.randn(B, device=DEV)
    for a, b in zip(K.kl_persample_bwd(mu, lv, g), S.kl_persample_bwd(mu, lv, g)):
        assert_f32_close(a, b, "kl bwd", 1e-6)


@pytest.mark.parametrize("B,n", [(8, 614400), (2, 1001), (1, 3)])
def test_mse(B, n):
    x, y = torch.rand(B, n, device=DEV), torch.rand(B, n, device=DEV)
    assert_f32_close(K.mse_persample_fwd(x, y), ((x.double() - y.double()) ** 2).sum(1).float(), "mse", 1e-5)
    g = torch.randn(B, device=DEV)
    dx, dy = K.mse_persample_bwd(x, y, g, True, True)
    rx, ry = S.mse_persample_bwd(x, y, g, True, True)
    assert torch.equal(dx, rx) and torch.equal(dy, ry)
    dx, dy = K.mse_persample_bwd(x, y, g, False, True)
    assert dx is None and torch.equal(dy, ry)


def test_layout_roundtrip():
    x = torch.randn(2, 24, 3, 4, 5, device=DEV)
    a = K.to_ndhwc_bf16(x)
    assert torch.equal(a, S.to_ndhwc_bf16(x))
    assert torch.equal(K.to_ncdhw_f32(a), x.to(torch.bfloat16).float())


def test_argument_errors_are_reported():
    with pytest.raises(K.SivaeError):
        K.conv3_igemm(bf(1, 4, 4, 4, 32), bf(27, 64, 32))            # Cin not a multiple of 64
    with pytest.raises(K.SivaeError):
        K.bn_act_fwd(bf(1, 3, 4, 4, 64), torch.ones(64, device=DEV), torch.zeros(64, device=DEV), None, 0.2, 1)  # odd D
    with pytest.raises(K.SivaeError):
        K.reparam_fwd(torch.zeros(4), torch.zeros(4), 0.1)             # CPU tensor


# ------------------------------------------------------------------------------------------- loss assembly
@pytest.mark.parametrize("B", [1, 8, 37])
def test_intro_loss_assembly_vs_torch_autograd(B):
    """Fused lossE / lossD assembly (utils/my_trainer.py:260-284, :301-321) and its gradients against the same
    expressions written with torch ops + autograd."""
    from sivae_b200 import functional as F
    torch.manual_seed(B)
    scale, b_rec, b_kl, b_neg, gamma_r = 8.0 / 614400, 1.0, 0.75, 1024.0, 1e-8

    def vec(lo, hi):
        return (torch.rand(B, device=DEV) * (hi - lo) + lo).requires_grad_(True)

    # E: magnitudes as at the start of training (r ~ 2e5, k ~ 2e3) -- the exp-ELBO terms then underflow towards 1e-26;
    # a second set with small values exercises the exponential branch at O(1)
    for (r_lo, r_hi, k_lo, k_hi) in ((1e5, 4e5, 5e2, 4e3), (1e2, 5e3, 1.0, 30.0)):
        vs = [vec(r_lo, r_hi), vec(k_lo, k_hi), vec(r_lo, r_hi), vec(k_lo, k_hi), vec(r_lo, r_hi), vec(k_lo, k_hi)]
        r_real, k_real, r_fake, k_fake, r_rec, k_rec = vs
        ef = (-2 * scale * (b_rec * r_fake + b_neg * k_fake)).exp().mean()
        er = (-2 * scale * (b_rec * r_rec + b_neg * k_rec)).exp().mean()
        ref = 10 * (scale * (b_rec * r_real.mean() + b_kl * k_real.mean()) + 0.5 * (ef + er))
        gref = torch.autograd.grad(ref, vs)
        vs2 = [v.detach().clone().requires_grad_(True) for v in vs]
        got, m_rr, m_kr, gef, ger = F.intro_loss_e(*vs2, scale, b_rec, b_kl, b_neg)
        ggot = torch.autograd.grad(got, vs2)
        assert float(got) == pytest.approx(float(ref), rel=2e-6)
        assert float(gef) == pytest.approx(float(ef), rel=1e-5, abs=1e-37) and float(ger) == pytest.approx(float(er), rel=1e-5, abs=1e-37)
        assert float(m_rr) == pytest.approx(float(r_real.mean()), rel=2e-6)
        for a, b in zip(ggot, gref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-30), float((a - b).abs().max())
    vs = [vec(1e5, 4e5), vec(5e2, 4e3), vec(5e2, 4e3), vec(1e5, 4e5), vec(1e5, 4e5)]
    r_real, k_rec, k_fake, r_rr, r_fr = vs
    ref = 10 * (scale * (b_rec * r_real.mean() + 0.5 * b_kl * (k_rec.mean() + k_fake.mean())
                         + gamma_r * 0.5 * b_rec * (r_rr.mean() + r_fr.mean())))
    gref = torch.autograd.grad(ref, vs)
    vs2 = [v.detach().clone().requires_grad_(True) for v in vs]
    outs = F.intro_loss_d(*vs2, scale, b_rec, b_kl, gamma_r)
    ggot = torch.autograd.grad(outs[0], vs2)
    assert float(outs[0]) == pytest.approx(float(ref), rel=2e-6)
    assert float(outs[2]) == pytest.approx(float(k_rec.mean()), rel=2e-6)
    for a, b in zip(ggot, gref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-30)


# ------------------------------------------------------------------------------------------- cluster-fused small BN
@pytest.mark.parametrize("shape", [(8, 10, 12, 10, 256), (2, 5, 6, 5, 128), (1, 4, 4, 4, 64), (3, 7, 9, 11, 256), (1, 2, 2, 4, 8)])
@pytest.mark.parametrize("with_res", [False, True])
def test_bn_cluster_small_tensors(shape, with_res):
    """sivae_bn_train_act_fwd / sivae_bn_act_bwd on small tensors can run as ONE launch of a 16-CTA cluster (partials
    through distributed shared memory; opt-in with SIVAE_BN_CLUSTER=1 because it measured slower in the step).  Both paths
    must agree with the specification, including running statistics, ragged voxel slices and the residual branch."""
    import os
    n, d, h, w, c = shape
    y, g = bf(*shape), bf(*shape)
    res = bf(*shape) if with_res else None
    gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV)

    def run():
        rm, rv, nbt = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
        out, mean, invstd = K.bn_train_act_fwd(y, res, gamma, beta, rm, rv, nbt, 0.1, 1e-5, 0.2)
        dconv, dres, dgam, dbet = K.bn_act_bwd(g, y, res, mean, invstd, gamma, beta, 0.2, 0, need_dres=with_res)
        return out, mean, invstd, rm, rv, int(nbt), dconv, dres, dgam, dbet

    old = os.environ.pop("SIVAE_BN_CLUSTER", None)
    try:
        b = run()                                    # default: statistics / finalize / apply kernels
        os.environ["SIVAE_BN_CLUSTER"] = "1"
        a = run()                                    # opt-in: single cluster launch
    finally:
        os.environ.pop("SIVAE_BN_CLUSTER", None)
        if old is not None:
            os.environ["SIVAE_BN_CLUSTER"] = old
    rm, rv, nbt = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    s_out, s_mean, s_invstd = S.bn_train_act_fwd(y, res, gamma, beta, rm, rv, nbt, 0.1, 1e-5, 0.2)
    s_dconv, s_dres, s_dgam, s_dbet = S.bn_act_bwd(g, y, res, s_mean, s_invstd, gamma, beta, 0.2, 0, need_dres=with_res)
    for got in (a, b):
        assert_bf16_close(got[0], s_out, "out")
        assert_f32_close(got[1], s_mean, "mean", 1e-5)
        assert_f32_close(got[2], s_invstd, "invstd", 1e-5)
        assert_f32_close(got[3], rm, "running_mean", 1e-5)
        assert_f32_close(got[4], rv, "running_var", 1e-5)
        assert got[5] == 1
        assert_bf16_close(got[6], s_dconv, "dconv", rel=2 ** -6)
        if with_res:
            assert_bf16_close(got[7], s_dres, "dres")
        assert_f32_close(got[8], s_dgam, "dgamma", 2e-4)
        assert_f32_close(got[9], s_dbet, "dbeta", 2e-4)
