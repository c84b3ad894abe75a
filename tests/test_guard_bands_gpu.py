"""Out-of-bounds WRITE detection without compute-sanitizer (the tool is closed on this GPU pool): every output and
workspace a C-ABI call allocates is placed between two 64 KiB guard bands filled with a sentinel byte; after the call
(and a device sync) the bands must be untouched.  Shapes are chosen so that tiles overhang the volume in every direction
(TMA stores must clip, SIMT kernels must guard) and so that every kernel family runs: tcgen05 convolutions (persistent
and tap-by-tap, fprop / dgrad / wgrad, upsample-folded), thin convolutions, BatchNorm passes incl. the keep-bit store,
loss / sampler kernels.

How: ``torch.empty`` / ``torch.empty_like`` / ``torch.zeros`` are wrapped for the duration of one wrapper call from
``sivae_b200.kernels`` (which allocates its outputs with exactly those), so the tensors the kernels write into are views
into guarded buffers.  OOB reads are not caught here; TMA zero-fills them by construction and the SIMT kernels' guards
are exercised by the parity tests on the same ragged shapes."""
import contextlib

import pytest
import torch

pytestmark = pytest.mark.gpu

import sivae_b200  # noqa: E402,F401
from sivae_b200 import kernels as K  # noqa: E402

DEV = "cuda"
GUARD = 64 * 1024
SENT = 0xA5


class Guards:
    def __init__(self):
        self.bufs = []

    def alloc(self, shape, dtype, device):
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 1024
        raw = _REAL["full"]((GUARD + nbytes + pad + GUARD,), SENT, dtype=torch.uint8, device=device)
        self.bufs.append((raw, nbytes, pad))
        return raw[GUARD:GUARD + nbytes].view(dtype).view(shape)

    def check(self, what):
        torch.cuda.synchronize()
        for raw, nbytes, pad in self.bufs:
            assert bool((raw[:GUARD] == SENT).all()), f"{what}: write BEFORE a {nbytes}-byte buffer"
            assert bool((raw[GUARD + nbytes:] == SENT).all()), f"{what}: write BEHIND a {nbytes}-byte buffer"


_REAL = {"empty": torch.empty, "empty_like": torch.empty_like, "zeros": torch.zeros, "full": torch.full}


def _norm_size(size):
    if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
        return tuple(size[0])
    return tuple(size)


@contextlib.contextmanager
def guarded(what):
    g = Guards()

    def empty(*size, dtype=None, device=None, **kw):
        if device is None or torch.device(device).type != "cuda":
            return _REAL["empty"](*size, dtype=dtype, device=device, **kw)
        return g.alloc(_norm_size(size), dtype or torch.float32, device)

    def empty_like(t, dtype=None, **kw):
        if not t.is_cuda:
            return _REAL["empty_like"](t, dtype=dtype, **kw)
        return g.alloc(tuple(t.shape), dtype or t.dtype, t.device)

    def zeros(*size, dtype=None, device=None, **kw):
        if device is None or torch.device(device).type != "cuda":
            return _REAL["zeros"](*size, dtype=dtype, device=device, **kw)
        return g.alloc(_norm_size(size), dtype or torch.float32, device).zero_()

    K._ws_cache.clear()                       # workspaces are allocated (guarded) inside the block too
    torch.empty, torch.empty_like, torch.zeros = empty, empty_like, zeros
    try:
        yield g
        torch.empty, torch.empty_like, torch.zeros = _REAL["empty"], _REAL["empty_like"], _REAL["zeros"]
        g.check(what)
    finally:
        torch.empty, torch.empty_like, torch.zeros = _REAL["empty"], _REAL["empty_like"], _REAL["zeros"]
        K._ws_cache.clear()


def bf(*shape):
    return torch.randn(*shape, device=DEV).to(torch.bfloat16)


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.manual_seed(3)
    yield
    torch.cuda.synchronize()


def test_the_guard_itself_detects_a_stray_write():
    with pytest.raises(AssertionError, match="BEHIND"):
        with guarded("self-test"):
            t = torch.empty(10, dtype=torch.float32, device=DEV)
            base = t.untyped_storage()
            torch.tensor([], dtype=torch.uint8, device=DEV).set_(base, GUARD + 40 + 2000, (4,)).fill_(0)


# (N, D, H, W): overhanging in w (8-wide / 16-wide tiles), h and d; one single-voxel case
RAGGED = [(2, 5, 19, 13), (1, 3, 17, 9), (1, 1, 1, 1), (3, 6, 30, 44)]


@pytest.mark.parametrize("shape", RAGGED)
@pytest.mark.parametrize("ci,co,env", [(64, 64, {}), (64, 64, {"SIVAE_CONV_KD": "force"}), (128, 64, {"SIVAE_CONV_KD": "force"}),
                                       (64, 128, {}), (256, 256, {}), (256, 256, {"SIVAE_N256": "4"})])
def test_conv_fprop_dgrad_wgrad_stay_in_bounds(shape, ci, co, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n, d, h, w = shape
    x, dy = bf(n, d, h, w, ci), bf(n, d, h, w, co)
    wt = torch.randn(co, ci, 3, 3, 3, device=DEV) * 0.05
    gamma, beta = torch.ones(co, device=DEV), torch.zeros(co, device=DEV)
    rm, rv = torch.zeros(co, device=DEV), torch.ones(co, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    with guarded(f"pack {shape} {ci}->{co}"):
        wf, wd = K.pack_conv3_weights(wt)
    with guarded(f"conv3_igemm {shape} {ci}->{co} {env}"):
        K.conv3_igemm(x, wf)
    with guarded(f"conv3_igemm_bn {shape} {ci}->{co} {env}"):
        K.conv3_igemm_bn(x, wf, gamma, beta, rm, rv, nbt, 0.1, 1e-5)
    with guarded(f"dgrad {shape} {co}->{ci} {env}"):
        K.conv3_igemm(dy, wd)
    with guarded(f"conv3_wgrad {shape} {ci}->{co}"):
        K.conv3_wgrad(x, dy)


@pytest.mark.parametrize("shape", [(2, 3, 9, 7), (1, 2, 16, 8), (1, 1, 1, 1), (2, 5, 20, 13)])
@pytest.mark.parametrize("ci,co,env", [(64, 64, {}), (64, 64, {"SIVAE_UPCONV_FUSED": "force", "SIVAE_UPWGRAD_TALL": "force"}),
                                       (128, 64, {}), (256, 128, {})])
def test_upsample_folded_convs_stay_in_bounds(shape, ci, co, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n, d, h, w = shape
    xlo, dyhi = bf(n, d, h, w, ci), bf(n, 2 * d, 2 * h, 2 * w, co)
    wt = torch.randn(co, ci, 3, 3, 3, device=DEV) * 0.05
    gamma, beta = torch.ones(co, device=DEV), torch.zeros(co, device=DEV)
    with guarded(f"pack_up {ci}->{co}"):
        wup, wupT = K.pack_upconv3_weights(wt)
    with guarded(f"upconv3_fprop {shape} {ci}->{co} {env}"):
        K.upconv3_fprop(xlo, wup)
    with guarded(f"upconv3_fprop_bn {shape} {ci}->{co} {env}"):
        K.upconv3_fprop_bn(xlo, wup, gamma, beta, None, None, None, 0.1, 1e-5)
    with guarded(f"upconv3_dgrad {shape} {ci}->{co}"):
        K.upconv3_dgrad(dyhi, wupT)
    with guarded(f"upconv3_wgrad {shape} {ci}->{co} {env}"):
        K.upconv3_wgrad(xlo, dyhi)


@pytest.mark.parametrize("shape", [(2, 5, 19, 13), (1, 1, 1, 1), (1, 4, 8, 16), (3, 2, 10, 6)])
def test_thin_convs_stay_in_bounds(shape):
    n, d, h, w = shape
    x1 = torch.rand(n, d, h, w, device=DEV)
    for c in (64, 128):
        a = bf(n, d, h, w, c)
        w27, w1 = torch.randn(c, 27, device=DEV) * 0.1, torch.randn(c, 1, device=DEV)
        b_c, b_1 = torch.randn(c, device=DEV), torch.randn(1, device=DEV)
        gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
        with guarded(f"c1_to_cn {shape} C={c}"):
            K.c1_to_cn(x1, w27, b_c)
            K.c1_to_cn(x1, w27, None, flip=True)
            K.c1_to_cn(x1, w1, b_c)
        with guarded(f"c1_to_cn_bn {shape} C={c}"):
            K.c1_to_cn_bn(x1, w27, b_c, gamma, beta, None, None, None, 0.1, 1e-5)
        with guarded(f"cn_to_c1 {shape} C={c}"):
            K.cn_to_c1(a, w27, b_1, False, 1, None, 0.35, 12345)
            K.cn_to_c1(a, w27, None, flip=True, act=0)
            K.cn_to_c1(a, w1, b_1)
        with guarded(f"wgrad_c1 {shape} C={c}"):
            K.wgrad_c1(a, x1, 27)
            K.wgrad_c1(a, x1, 27, flip=True)
            K.wgrad_c1(a, x1, 1)


@pytest.mark.parametrize("shape", [(2, 4, 6, 10, 64), (1, 1, 1, 1, 128), (3, 2, 2, 2, 256), (1, 6, 10, 14, 64)])
def test_batchnorm_passes_stay_in_bounds(shape):
    c = shape[-1]
    y, g = bf(*shape), bf(*shape)
    gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.2
    with guarded(f"bn_train_coeffs {shape}"):
        mean, invstd, scale, shift = K.bn_train_coeffs(y, gamma, beta, None, None, None, 0.1, 1e-5)
    even = all(s % 2 == 0 for s in shape[1:4])
    with guarded(f"bn_act_fwd {shape}"):
        K.bn_act_fwd(y, scale, shift, None, 0.2, 0)
        K.bn_act_fwd(y, scale, shift, y, 0.2, 0)
        K.bn_act_fwd(y, scale, shift, None, 0.2, 2)
        if even:
            K.bn_act_fwd(y, scale, shift, None, 0.2, 1)
    with guarded(f"bn_train_act_fwd {shape}"):
        K.bn_train_act_fwd(y, None, gamma, beta, None, None, None, 0.1, 1e-5, 0.2)
    with guarded(f"keep-bit store {shape}"):
        bits = torch.empty(y.numel() // 8, dtype=torch.uint8, device=DEV)
        K.bn_act_fwd(y, scale, shift, None, 0.2, 0, None, 0.35, 4242, keep_bits=bits)
        K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0, None, 0.35, 4242, keep_bits=bits)
    with guarded(f"bn_act_bwd {shape}"):
        K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0)
        K.bn_act_bwd(g, y, y, mean, invstd, gamma, beta, 0.2, 0, need_dres=True)
        K.bn_act_bwd(g, y, None, mean, invstd, gamma, beta, 0.2, 0, None, 0.35, 99)
        if even:
            gp = bf(shape[0], shape[1] // 2, shape[2] // 2, shape[3] // 2, c)
            K.bn_act_bwd(gp, y, None, mean, invstd, gamma, beta, 0.2, 1)


@pytest.mark.parametrize("B,n", [(1, 1), (3, 1201), (8, 1200), (2, 4099)])
def test_latent_and_loss_kernels_stay_in_bounds(B, n):
    mu, lv, eps = (torch.randn(B, n, device=DEV) for _ in range(3))
    x, yv, gvec = torch.rand(B, n, device=DEV), torch.rand(B, n, device=DEV), torch.rand(B, device=DEV)
    with guarded(f"reparam {B}x{n}"):
        z = K.reparam_fwd(mu, lv, eps)
        K.reparam_draw_fwd(mu, lv, 777)
        K.reparam_bwd(z, lv, eps)
    with guarded(f"kl {B}x{n}"):
        K.kl_persample_fwd(mu, lv)
        K.kl_persample_bwd(mu, lv, gvec)
    with guarded(f"mse {B}x{n}"):
        K.mse_persample_fwd(x, yv)
        K.mse_persample_bwd(x, yv, gvec, True, True)
    with guarded(f"intro loss B={B}"):
        vs = [torch.rand(B, device=DEV) * 100 for _ in range(6)]
        K.intro_loss_e_fwd(*vs, 1e-5, 1.0, 0.75, 1024.0)
        K.intro_loss_d_fwd(*vs[:5], 1e-5, 1.0, 0.75, 1e-8)
