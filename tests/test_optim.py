"""optim.FusedAdam (SURVEY section 8f NEXT-2) against torch.optim.Adam, the optimiser the reference constructs in
utils/my_trainer.py:183-184: identical trajectories, state layout, skipping of gradient-less parameters, LR-scheduler
support, and the in-kernel refresh of the bf16 convolution weight packs.
CPU: host logic with libsivae.so replaced by the kernel specification.  GPU: the CUDA kernel through the C ABI."""
import pytest
import torch

import sivae_b200
from sivae_b200 import functional as F
from sivae_b200 import kernels as K
from tests.emu import emulated_kernels


def _run(device, steps=5):
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(64, 64, 3, padding=1, bias=False).to(device)
    lin = torch.nn.Linear(7, 5).to(device)
    unused = torch.nn.Linear(3, 3).to(device)                     # never gets a gradient
    params = list(conv.parameters()) + list(lin.parameters()) + list(unused.parameters())
    ref_params = [p.detach().clone().requires_grad_(True) for p in params]
    opt = sivae_b200.FusedAdam(params, lr=2e-4)
    ref = torch.optim.Adam(ref_params, lr=2e-4)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=(3,), gamma=0.1)
    sch_ref = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=(3,), gamma=0.1)
    w = conv.weight
    wf, wd = F._packed(w)                                          # packs exist before the first update
    for step in range(steps):
        for plist in (params, ref_params):
            for p in plist:
                p.grad = None
        g = torch.Generator(device="cpu").manual_seed(100 + step)
        for p, q in zip(params[:3], ref_params[:3]):
            gr = torch.randn(p.shape, generator=g).to(device) * (10.0 ** (step - 2))
            p.grad = gr.clone()
            q.grad = gr.clone()
        opt.step()
        ref.step()
        sch.step()
        sch_ref.step()
        for p, q in zip(params, ref_params):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-9), (step, float((p - q).abs().max()))
        # the packs of the convolution weight were refreshed in place and are still the cached, current ones
        packs = F.current_packs(w, up=False)
        assert packs is not None and packs[0] is wf and packs[1] is wd
        wf2, wd2 = K.pack_conv3_weights(w.detach())              # a fresh pack of the updated weight
        assert torch.equal(wf, wf2) and torch.equal(wd, wd2)
        if wf.dtype == torch.bfloat16:                             # and the layout itself: wf[tap][co][ci], wd[26-tap][ci][co]
            wq = w.detach().to(torch.bfloat16)
            assert torch.equal(wf, wq.permute(2, 3, 4, 0, 1).reshape(27, 64, 64))
            assert torch.equal(wd, wq.flip(2, 3, 4).permute(2, 3, 4, 1, 0).reshape(27, 64, 64))
    assert all(p.grad is None for p in unused.parameters()) and all(len(opt.state[p]) == 0 for p in unused.parameters())
    st = opt.state[w]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and int(st["step"]) == steps
    # moments: same formulas, fp32 rounding order may differ (fma contraction); gradients here span 1e-2 .. 1e2
    m_ref, v_ref = ref.state[ref_params[0]]["exp_avg"], ref.state[ref_params[0]]["exp_avg_sq"]
    assert torch.allclose(st["exp_avg"], m_ref, rtol=1e-4, atol=1e-5 * float(m_ref.abs().max()))
    assert torch.allclose(st["exp_avg_sq"], v_ref, rtol=1e-4, atol=1e-6 * float(v_ref.abs().max()))
    sd = opt.state_dict()                                          # serialisable like a torch optimiser
    assert len(sd["param_groups"][0]["params"]) == len(params)


def test_fused_adam_matches_torch_cpu_emulated():
    with emulated_kernels():
        _run("cpu")


def test_stale_upconv_packs_are_dropped_cpu_emulated():
    with emulated_kernels():
        conv = torch.nn.Conv3d(64, 64, 3, padding=1, bias=False)
        w = conv.weight
        up = F._packed(w, True)
        assert F.current_packs(w, up=True) is up
        w.grad = torch.ones_like(w)
        sivae_b200.FusedAdam([w], lr=1e-3).step()
        assert F.current_packs(w, up=True) is None                 # not refreshed in-kernel -> must be re-packed
        assert not torch.equal(F._packed(w, True)[0], up[0])


def _resume(device):
    """state_dict -> fresh optimiser -> load_state_dict continues the trajectory of an uninterrupted torch.optim.Adam
    (the shared device step counter and lr are persisted and re-bound), and a parameter whose first gradient arrives
    after the group's first step is refused (one bias-correction counter per group)."""
    torch.manual_seed(0)
    ps = [torch.randn(n, device=device, requires_grad=True) for n in (5, 33)]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    opt, ref = sivae_b200.FusedAdam(ps, lr=1e-3), torch.optim.Adam(qs, lr=1e-3)

    def one(o, r, step):
        g = torch.Generator(device="cpu").manual_seed(step)
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, generator=g).to(device)
            p.grad, q.grad = gr.clone(), gr.clone()
        o.step()
        r.step()

    for step in range(3):
        one(opt, ref, step)
    sd = opt.state_dict()
    opt2 = sivae_b200.FusedAdam(ps, lr=5e-2)                      # lr comes from the checkpoint, not the ctor
    opt2.load_state_dict(sd)
    assert opt2.state[ps[0]]["step"] is opt2.state[ps[1]]["step"] and int(opt2.state[ps[0]]["step"]) == 3
    for step in range(3, 6):
        one(opt2, ref, step)
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, rtol=2e-6, atol=1e-8)
    assert int(opt2.state[ps[0]]["step"]) == 6
    late = torch.randn(4, device=device, requires_grad=True)
    opt3 = sivae_b200.FusedAdam([ps[0], late], lr=1e-3)
    ps[0].grad = torch.ones_like(ps[0])
    opt3.step()
    late.grad = torch.ones_like(late)
    with pytest.raises(RuntimeError, match="first gradient after"):
        opt3.step()


def test_fused_adam_resume_cpu_emulated():
    with emulated_kernels():
        _resume("cpu")


@pytest.mark.gpu
def test_fused_adam_resume_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    _resume("cuda")


@pytest.mark.gpu
def test_fused_adam_matches_torch_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    _run("cuda")


@pytest.mark.gpu
def test_fused_adam_many_tensors_gpu():
    """More tensors than one launch carries (32), odd sizes, against torch.optim.Adam."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.manual_seed(1)
    params = [torch.randn(n, device="cuda", requires_grad=True) for n in [1, 3, 17, 1000, 4097] * 15]
    ref_params = [p.detach().clone().requires_grad_(True) for p in params]
    opt, ref = sivae_b200.FusedAdam(params, lr=1e-3), torch.optim.Adam(ref_params, lr=1e-3)
    for step in range(3):
        for p, q in zip(params, ref_params):
            p.grad = torch.randn_like(p)
            q.grad = p.grad.clone()
        opt.step()
        ref.step()
    for p, q in zip(params, ref_params):
        assert torch.allclose(p, q, rtol=2e-6, atol=1e-8)
