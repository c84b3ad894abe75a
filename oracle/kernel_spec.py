"""Executable specification of every function in ``soft-intro-vae-for-3d-mri_b200/kernels.py``
(TEST INFRASTRUCTURE ONLY -- plain torch, device-agnostic, never on the product path).

Two uses:
  * ``-m gpu`` tests compare each CUDA kernel (called through the C ABI) with the function of the
    same name here, on the same inputs;
  * ``-m "not gpu"`` tests monkeypatch these functions over ``kernels.*`` so the whole autograd
    wiring of the drop-in modules can be checked against ``oracle/sivae_oracle.py`` on a CPU, in
    fp32 ("exact" mode: activations keep whatever dtype they come in with; the CUDA path uses bf16).

Math is done in fp32 and rounded to the activation dtype exactly where the kernels round.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

RESAMPLE_NONE, RESAMPLE_AVGPOOL2, RESAMPLE_UPSAMPLE2 = 0, 1, 2

# dtype of activations created from fp32 inputs (c1_to_cn).  bf16 mirrors the CUDA path; the CPU
# wiring tests set this to float32 to separate wiring errors from rounding.
ACT_DTYPE = torch.bfloat16


def _ncdhw(x):  # [N,D,H,W,C] -> [N,C,D,H,W] fp32
    return x.float().permute(0, 4, 1, 2, 3)


def _ndhwc(x, dtype):  # [N,C,D,H,W] -> [N,D,H,W,C]
    return x.permute(0, 2, 3, 4, 1).contiguous().to(dtype)


def launch_count():
    return 0


def device_check():
    return None


def set_seed_counter(counter):
    return None


def advance_seed_counter(device):
    return None


def pack_conv3_weights(w):
    co, ci = w.shape[:2]
    w27 = w.reshape(co, ci, 27)
    wf = w27.permute(2, 0, 1).contiguous().to(ACT_DTYPE)                 # [27,Co,Ci]
    wd = w27.flip(2).permute(2, 1, 0).contiguous().to(ACT_DTYPE)         # [27,Ci,Co], tap flipped
    return wf, wd


def conv3_igemm(x, wpack):
    co, ci = wpack.shape[1], wpack.shape[2]
    w = wpack.float().permute(1, 2, 0).reshape(co, ci, 3, 3, 3)
    y = F.conv3d(_ncdhw(x), w, None, 1, 1)
    return _ndhwc(y, x.dtype)


def conv3_igemm_bn(x, wpack, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps):
    y = conv3_igemm(x, wpack)
    return (y,) + tuple(bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps))


def conv3_wgrad(x, dy):
    xi, g = _ncdhw(x), _ncdhw(dy)
    co, ci = g.shape[1], xi.shape[1]
    return torch.nn.grad.conv3d_weight(xi, (co, ci, 3, 3, 3), g, stride=1, padding=1)


def _up_tapset(p, a):
    """3x3x3 taps (per axis) that land on low-res offset index a for output parity p."""
    return ([0], [1, 2])[a] if p == 0 else ([0, 1], [2])[a]


def pack_upconv3_weights(w):
    co, ci = w.shape[:2]
    wup = torch.zeros(64, co, ci, dtype=torch.float32, device=w.device)
    for p in range(8):
        pd, ph, pw = p >> 2, (p >> 1) & 1, p & 1
        for abc in range(8):
            a, b, c = abc >> 2, (abc >> 1) & 1, abc & 1
            acc = 0
            for kd in _up_tapset(pd, a):
                for kh in _up_tapset(ph, b):
                    for kw in _up_tapset(pw, c):
                        acc = acc + w[:, :, kd, kh, kw]
            wup[p * 8 + abc] = acc
    return wup.to(ACT_DTYPE), wup.transpose(1, 2).contiguous().to(ACT_DTYPE)


def _upconv3_fprop_f32(x_lo, wup):
    n, d, h, w, ci = x_lo.shape
    co = wup.shape[1]
    xp = F.pad(x_lo.float(), (0, 0, 1, 1, 1, 1, 1, 1))
    y = torch.zeros(n, 2 * d, 2 * h, 2 * w, co, dtype=torch.float32, device=x_lo.device)
    for p in range(8):
        pd, ph, pw = p >> 2, (p >> 1) & 1, p & 1
        acc = 0
        for abc in range(8):
            a, b, c = abc >> 2, (abc >> 1) & 1, abc & 1
            od, oh, ow = a - 1 + pd, b - 1 + ph, c - 1 + pw
            xs = xp[:, 1 + od:1 + od + d, 1 + oh:1 + oh + h, 1 + ow:1 + ow + w]
            acc = acc + xs @ wup[p * 8 + abc].float().t()
        y[:, pd::2, ph::2, pw::2] = acc
    return y


def upconv3_fprop(x_lo, wup):
    return _upconv3_fprop_f32(x_lo, wup).to(x_lo.dtype)


def upconv3_fprop_bn(x_lo, wup, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps):
    y = upconv3_fprop(x_lo, wup)
    return (y,) + tuple(bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps))


def upconv3_dgrad(dy_hi, wupT):
    n, d2, h2, w2, co = dy_hi.shape
    ci = wupT.shape[1]
    with torch.enable_grad():
        x = torch.zeros(n, d2 // 2, h2 // 2, w2 // 2, ci, dtype=torch.float32, device=dy_hi.device, requires_grad=True)
        y = _upconv3_fprop_f32(x, wupT.transpose(1, 2))
        (g,) = torch.autograd.grad(y, x, dy_hi.float())
    return g.to(dy_hi.dtype)


def upconv3_wgrad(x_lo, dy_hi):
    xu = x_lo.float().repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)
    xi, g = xu.permute(0, 4, 1, 2, 3), dy_hi.float().permute(0, 4, 1, 2, 3)
    return torch.nn.grad.conv3d_weight(xi, (g.shape[1], xi.shape[1], 3, 3, 3), g, stride=1, padding=1)


def bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps):
    c = y.shape[-1]
    f = y.float().reshape(-1, c)
    n = f.shape[0]
    mean = f.mean(0)
    var = f.var(0, unbiased=False)
    invstd = torch.rsqrt(var + eps)
    scale = gamma * invstd
    shift = beta - mean * scale
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
    if running_var is not None:
        running_var.mul_(1 - momentum).add_(momentum * var * (n / max(n - 1, 1)))
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    return mean, invstd, scale, shift


def _unpack_keep_bits(bits, shape):
    """uint8 [numel/8] (bit k of byte i = element 8i+k kept) -> bool tensor of ``shape``."""
    k = torch.arange(8, device=bits.device, dtype=torch.uint8)
    return ((bits.reshape(-1, 1) >> k) & 1).bool().reshape(shape)


def _keep_scale(y, mask, p, keep_bits=None, draw=False):
    """Keep-scale of a dropout call.  ``mask``: explicit byte keep-mask.  ``keep_bits``: the keep-bit store -- the forward
    (``draw=True``) draws the decisions and writes them there (here from torch's generator: the Philox stream itself is
    only checked statistically), the backward reads them back."""
    if mask is not None:
        return mask.float() * (1.0 / (1.0 - p))
    if keep_bits is not None and p > 0.0:
        if draw:
            keep = torch.rand(y.shape, device=y.device) >= p
            k = torch.arange(8, device=y.device)
            keep_bits.copy_((keep.reshape(-1, 8).long() << k).sum(1).to(torch.uint8))
        return _unpack_keep_bits(keep_bits, y.shape).float() * (1.0 / (1.0 - p))
    if p > 0.0:
        raise NotImplementedError("Philox dropout is only checked statistically; pass an explicit mask")
    return None


def _resample(a, resample):  # a: [N,D,H,W,C] fp32
    if resample == RESAMPLE_AVGPOOL2:
        return F.avg_pool3d(a.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1)
    if resample == RESAMPLE_UPSAMPLE2:
        return a.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)
    return a


def _resample_T(g, resample):  # transpose of _resample; g fp32 in output space
    if resample == RESAMPLE_AVGPOOL2:
        return g.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.125
    if resample == RESAMPLE_UPSAMPLE2:
        return F.avg_pool3d(g.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1) * 8.0
    return g


def bn_act_fwd(y, scale, shift, res, slope, resample, mask=None, p=0.0, seed=0, keep_bits=None):
    t = y.float() * scale + shift
    if res is not None:
        t = t + res.float()
    a = torch.where(t > 0, t, slope * t)
    ks = _keep_scale(y, mask, p, keep_bits, draw=True)
    if ks is not None:
        a = a * ks
    return _resample(a, resample).contiguous().to(y.dtype)


SMALL_BN_ELEMS = 3 << 20


def bn_train_act_fwd(y, res, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, slope):
    mean, invstd, scale, shift = bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps)
    return bn_act_fwd(y, scale, shift, res, slope, 0), mean, invstd


def bn_act_bwd(g, y, res, mean, invstd, gamma, beta, slope, resample, mask=None, p=0.0, seed=0,
               need_dres=False, need_affine=True, keep_bits=None):
    c = y.shape[-1]
    xh = (y.float() - mean) * invstd
    t = xh * gamma + beta
    if res is not None:
        t = t + res.float()
    gp = _resample_T(g.float(), resample)
    ks = _keep_scale(y, mask, p, keep_bits)
    if ks is not None:
        gp = gp * ks
    dt = gp * torch.where(t > 0, torch.ones_like(t), torch.full_like(t, slope))
    n = y.numel() // c
    s1 = dt.reshape(-1, c).sum(0)
    s2 = (dt * xh).reshape(-1, c).sum(0)
    dconv = gamma * invstd * (dt - s1 / n - xh * (s2 / n))
    return (dconv.to(y.dtype), dt.to(y.dtype) if need_dres else None,
            s2 if need_affine else None, s1 if need_affine else None)


def _w5(w, flip):  # [C,T] -> conv kernel taps [C,k,k,k]
    c, t = w.shape
    k = 3 if t == 27 else 1
    w = w.flip(1) if flip else w
    return w.reshape(c, k, k, k), k


def c1_to_cn(x1, w, bias, flip=False, out=None):
    w5, k = _w5(w, flip)
    y = F.conv3d(x1.float().unsqueeze(1), w5.unsqueeze(1), bias, 1, k // 2)      # [N,C,D,H,W]
    y = y.permute(0, 2, 3, 4, 1)
    if out is not None:
        out.copy_((out.float() + y).to(out.dtype))
        return out
    return y.contiguous().to(ACT_DTYPE)


def c1_to_cn_bn(x1, w, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps):
    y = c1_to_cn(x1, w, bias)
    return (y,) + tuple(bn_train_coeffs(y, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps))


def cn_to_c1(x, w, bias, flip=False, act=0, mask=None, p=0.0, seed=0):
    w5, k = _w5(w, flip)
    y = F.conv3d(_ncdhw(x), w5.unsqueeze(0), bias, 1, k // 2)[:, 0]             # [N,D,H,W]
    if act == 1:
        y = F.relu(y)
        if mask is not None:
            y = y * mask.float() * (1.0 / (1.0 - p))
        elif p > 0.0:
            raise NotImplementedError("Philox dropout is only checked statistically; pass an explicit mask")
    return y.contiguous()


def wgrad_c1(xc, x1, taps, flip=False):
    k = 3 if taps == 27 else 1
    xi = x1.float().unsqueeze(1)                                                   # [N,1,D,H,W]
    gc = _ncdhw(xc)                                                                # [N,C,D,H,W]
    c = gc.shape[1]
    # dw[c,t] = sum_v xc[v,c] * x1[v+delta(t)]  == weight gradient of conv3d(x1 -> C) with output grad xc
    dw = torch.nn.grad.conv3d_weight(xi, (c, 1, k, k, k), gc, stride=1, padding=k // 2).reshape(c, taps)
    if flip:
        dw = dw.flip(1)
    return dw.contiguous(), gc.sum((0, 2, 3, 4)), x1.float().sum().reshape(1)


def relu_drop_bwd(g, out, p):
    return torch.where(out > 0, g * (1.0 / (1.0 - p)), torch.zeros_like(g))


def reparam_fwd(mu, logvar, eps):
    return mu + eps * torch.exp(0.5 * logvar)


def reparam_draw_fwd(mu, logvar, seed):
    """eps ~ N(0,1) (here from torch's generator: the Philox stream itself is only checked statistically) -> (z, eps)."""
    eps = torch.randn_like(mu)
    return reparam_fwd(mu, logvar, eps), eps


def reparam_bwd(dz, logvar, eps):
    return dz.clone(), dz * eps * torch.exp(0.5 * logvar) * 0.5


def kl_persample_fwd(mu, logvar):
    return -0.5 * torch.sum(1 + logvar - mu * mu - logvar.exp(), dim=1)


def kl_persample_bwd(mu, logvar, g):
    return g[:, None] * mu, g[:, None] * 0.5 * (logvar.exp() - 1.0)


def mse_persample_fwd(x, y):
    return ((x - y) ** 2).sum(1)


def mse_persample_bwd(x, y, g, need_dx, need_dy):
    d = 2.0 * (x - y) * g[:, None]
    return (d if need_dx else None), (-d if need_dy else None)


def to_ndhwc_bf16(x):
    return _ndhwc(x, ACT_DTYPE)


def to_ncdhw_f32(x):
    return _ncdhw(x).contiguous()


def intro_loss_e_fwd(r_real, k_real, r_fake, k_fake, r_rec, k_rec, scale, b_rec, b_kl, b_neg):
    """utils/my_trainer.py:260-284 on per-sample vectors."""
    ef = (-2 * scale * (b_rec * r_fake + b_neg * k_fake)).exp().mean()
    er = (-2 * scale * (b_rec * r_rec + b_neg * k_rec)).exp().mean()
    m_rr, m_kr = r_real.mean(), k_real.mean()
    loss = 10 * (scale * (b_rec * m_rr + b_kl * m_kr) + 0.5 * (ef + er))
    return torch.stack([loss, m_rr, m_kr, ef, er]).float()


def intro_loss_e_bwd(r_fake, k_fake, r_rec, k_rec, g, scale, b_rec, b_kl, b_neg):
    b = r_fake.numel()
    go = 10.0 * g.reshape(-1)[0] / b
    ef = (-2 * scale * (b_rec * r_fake + b_neg * k_fake)).exp()
    er = (-2 * scale * (b_rec * r_rec + b_neg * k_rec)).exp()
    one = torch.ones(b, dtype=torch.float32, device=r_fake.device)
    return torch.stack([go * scale * b_rec * one, go * scale * b_kl * one, go * 0.5 * ef * (-2 * scale * b_rec),
                        go * 0.5 * ef * (-2 * scale * b_neg), go * 0.5 * er * (-2 * scale * b_rec),
                        go * 0.5 * er * (-2 * scale * b_neg)]).float()


def intro_loss_d_fwd(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, scale, b_rec, b_kl, gamma_r):
    """utils/my_trainer.py:301-321 on per-sample vectors."""
    m = [t.mean() for t in (r_real, k_rec, k_fake, r_rec_rec, r_fake_rec)]
    loss = 10 * (scale * (b_rec * m[0] + 0.5 * b_kl * (m[1] + m[2]) + gamma_r * 0.5 * b_rec * (m[3] + m[4])))
    return torch.stack([loss] + m).float()


def intro_loss_d_bwd(g, batch, scale, b_rec, b_kl, gamma_r):
    go = 10.0 * g.reshape(-1)[0] * scale / batch
    one = torch.ones(batch, dtype=torch.float32, device=g.device)
    return torch.stack([go * b_rec * one, go * 0.5 * b_kl * one, go * 0.5 * b_kl * one, go * gamma_r * 0.5 * b_rec * one,
                        go * gamma_r * 0.5 * b_rec * one]).float()


def volume_stats(x):
    v = x.reshape(x.shape[0], -1).double()
    return torch.stack([v.mean(1), v.std(1, unbiased=False), v.min(1).values, v.max(1).values], 1).float()


def preprocess_clip_minmax(x, cut_range=4.0, out=None):
    """utils/data_load.py:25-30: np.clip(v, 0, 4*np.std(v)); (v - min) / (max - min), per volume."""
    stats = volume_stats(x)
    v = x.reshape(x.shape[0], -1)
    c = torch.minimum(v.clamp_min(0.0), cut_range * stats[:, 1:2])
    lo, hi = c.min(1, keepdim=True).values, c.max(1, keepdim=True).values
    y = ((c - lo) / (hi - lo)).reshape(x.shape)
    if out is not None:
        out.copy_(y)
        y = out
    return y, stats


def affine_resample(x, mats, pad=None, stats=None):
    """Trilinear resampling through output->input voxel matrices, via F.grid_sample(align_corners=True); samples outside
    the volume take the pad value (pad[b], else stats[b][2] = per-volume minimum, else 0)."""
    b, d, h, w = x.shape
    dev = x.device
    dd, hh, ww = torch.meshgrid(torch.arange(d, device=dev), torch.arange(h, device=dev), torch.arange(w, device=dev),
                                indexing="ij")
    ones = torch.ones_like(dd)
    out_idx = torch.stack([dd, hh, ww, ones], -1).double().reshape(-1, 4)            # [n,4]
    m = mats.double().reshape(b, 3, 4)
    src = torch.einsum("bij,nj->bni", m, out_idx)                                    # [b,n,3] = (d,h,w) input coords
    size = torch.tensor([d, h, w], device=dev, dtype=torch.float64)
    norm = 2.0 * src / (size - 1).clamp_min(1) - 1.0
    grid = norm[..., [2, 1, 0]].reshape(b, d, h, w, 3).float()                       # grid_sample wants (x=w, y=h, z=d)
    padv = pad if pad is not None else (stats[:, 2] if stats is not None else torch.zeros(b, device=dev))
    padv = padv.reshape(b, 1, 1, 1, 1).float()
    y = F.grid_sample(x[:, None] - padv, grid, mode="bilinear", padding_mode="zeros", align_corners=True) + padv
    return y[:, 0]


def linear_fwd(x, weight, bias, relu=False):
    """nn.Linear (+ ReLU): mymodel.py:125,:150-153."""
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


def linear_dgrad(dy, weight):
    return dy @ weight


def linear_wgrad(x, dy, need_bias=True):
    return dy.t() @ x, (dy.sum(0) if need_bias else None)


def ndhwc_to_flat(h, c, gate=None):
    """NDHWC [B,d,h,w,Cp] -> fp32 [B, c*S] in NCDHW flatten order (mymodel.py:140), optionally gated by gate > 0."""
    b = h.shape[0]
    out = h[..., :c].permute(0, 4, 1, 2, 3).reshape(b, -1).float()
    return out if gate is None else out * (gate > 0).float()


def flat_to_ndhwc(y, c, cp, grid):
    """fp32 [B, c*S] -> NDHWC [B,*grid,cp] in the activation dtype, channels zero-padded (mymodel.py:219)."""
    b = y.shape[0]
    v = y.reshape(b, c, *grid).permute(0, 2, 3, 4, 1)
    if cp != c:
        v = torch.cat([v, torch.zeros(*v.shape[:-1], cp - c, dtype=v.dtype, device=v.device)], dim=-1)
    return v.contiguous().to(ACT_DTYPE)


def add_act_fwd(a, b, slope):
    return F.leaky_relu(a.float() + b.float(), slope).to(ACT_DTYPE)


def add_act_bwd(g, out, slope):
    return torch.where(out.float() > 0, g.float(), g.float() * slope).to(ACT_DTYPE)


def similarity_topk(queries, database, k, metric="cosine"):
    """Top-k similarity (cosine, or negative squared L2), best first, ties -> lower database index."""
    q, d = queries.double(), database.double()
    dot = q @ d.t()
    qn, dn = (q * q).sum(1, keepdim=True), (d * d).sum(1, keepdim=True).t()
    s = dot / (qn * dn).clamp_min(1e-30).sqrt() if metric == "cosine" else -(qn + dn - 2 * dot)
    order = torch.argsort(s, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(s, 1, order).float(), order.to(torch.int32)


def adam_step(tensors, lr, beta1, beta2, eps, step):
    """torch/optim/adam.py _single_tensor_adam (no weight decay / amsgrad / maximize), which is what the reference's
    torch.optim.Adam(lr=2e-4) runs (utils/my_trainer.py:183-184,:288,:324); refreshes the given bf16 packs."""
    t = int(step.item()) + 1
    lr_v = float(lr.item())
    bc1 = 1.0 - beta1 ** t
    bc2_sqrt = (1.0 - beta2 ** t) ** 0.5
    for p, g, m, v, packs in tensors:
        m.lerp_(g, 1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.data.addcdiv_(m, denom, value=-(lr_v / bc1))   # .data: like the kernel, no version-counter bump
        if packs is not None:
            wf, wd = pack_conv3_weights(p.detach())
            packs[0].copy_(wf)
            if packs[1] is not None:
                packs[1].copy_(wd)
    step += 1


ALL = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("F",)]
