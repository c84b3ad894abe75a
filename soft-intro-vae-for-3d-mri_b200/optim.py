"""Fused optimiser of the training hot path (NEXT-2).

``FusedAdam`` is a drop-in for the ``torch.optim.Adam(params, lr=2e-4)`` objects the reference builds in
``utils/my_trainer.py:183-184`` and steps at ``:288`` / ``:324``: same constructor arguments that the reference uses
(``lr``, ``betas``, ``eps``), same update rule and operation order, same per-parameter ``state`` keys
(``step``, ``exp_avg``, ``exp_avg_sq``), parameters whose ``grad is None`` are skipped, and
``param_groups[i]['lr']`` is honoured so ``MultiStepLR`` (:185-186) keeps working.

One ``step()`` is one C-ABI call (``sivae_adam_step``): a multi-tensor kernel updates every parameter of the group and,
for 3x3x3 convolution weights, rewrites their bf16 tap-major packs (forward + data-gradient layouts) in the same pass --
replacing torch's ~10 foreach launches per group and one ``pack_conv3_weights`` launch per convolution.  ``lr`` and the
step counter are device scalars, so the call is CUDA-graph capturable and replays stay correct.

Difference from ``torch.optim.Adam``: the step counter (bias correction) is ONE device scalar per parameter group, shared
by every parameter of the group (``state[p]['step']`` is that tensor).  torch gives each parameter its own counter that
starts when the parameter first receives a gradient; here a parameter that joins later would inherit the group's count,
so ``step()`` refuses it (the hot path never does this: the set of gradient-bearing parameters of a phase is fixed,
SURVEY Q1/Q2).  ``state_dict()`` / ``load_state_dict()`` persist and re-bind the shared counter and the device ``lr``.
"""
from __future__ import annotations

import torch

from . import functional as F
from . import kernels as K


class FusedAdam(torch.optim.Optimizer):
    graph_safe = True       # graph.GraphedTrainStep accepts it (no host-side state changes per step)

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self._gs = {}           # group index -> dict(step=int64[1], lr=fp32[1], lr_host=float)

    def _group_scalars(self, gi, group, device):
        gs = self._gs.get(gi)
        if gs is None or gs["step"].device != device:
            gs = dict(step=torch.zeros(1, dtype=torch.int64, device=device),
                      lr=torch.full((1,), float(group["lr"]), dtype=torch.float32, device=device),
                      lr_host=float(group["lr"]), started=False)
            self._gs[gi] = gs
        return gs

    def sync_lr(self):
        """Push ``param_groups[i]['lr']`` to the device scalars (call after an LR-scheduler step when the optimiser
        step itself is replayed from a CUDA graph; ``step()`` does it automatically in eager mode)."""
        for gi, group in enumerate(self.param_groups):
            gs = self._gs.get(gi)
            if gs is not None and gs["lr_host"] != float(group["lr"]):
                gs["lr"].fill_(float(group["lr"]))
                gs["lr_host"] = float(group["lr"])

    def load_state_dict(self, state_dict):
        """torch deep-copies every state tensor on load, which would leave each parameter with a private ``step``
        nobody advances: re-bind the group's shared device counter (value from the checkpoint) and the device ``lr``."""
        super().load_state_dict(state_dict)
        self._gs = {}
        for gi, group in enumerate(self.param_groups):
            have = [p for p in group["params"] if len(self.state.get(p, {})) != 0]
            if not have:
                continue
            gs = self._group_scalars(gi, group, have[0].device)
            gs["step"].copy_(torch.as_tensor(self.state[have[0]]["step"]).reshape(1).to(torch.int64))
            gs["started"] = True
            for p in have:
                self.state[p]["step"] = gs["step"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            gs = self._group_scalars(gi, group, params[0].device)
            if gs["lr_host"] != float(group["lr"]):
                gs["lr"].fill_(float(group["lr"]))
                gs["lr_host"] = float(group["lr"])
            b1, b2 = group["betas"]
            tensors = []
            for p in params:
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    if gs["started"]:
                        raise RuntimeError("FusedAdam: a parameter received its first gradient after the group's first "
                                           "step; the group shares one bias-correction counter (see module docstring)")
                    st["step"] = gs["step"]
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                packs = F.current_packs(p, up=False) if p.dim() == 5 else None
                tensors.append((p, g, st["exp_avg"], st["exp_avg_sq"], packs))
            K.adam_step(tensors, gs["lr"], float(b1), float(b2), float(group["eps"]), gs["step"])
            gs["started"] = True
            # the update bypasses torch's version counter: forget every cached pack that was not refreshed in-kernel
            for p, _, _, _, packs in tensors:
                if p.dim() == 5:
                    F.drop_packs(p, up=True)
                    if packs is None:
                        F.drop_packs(p, up=False)
        return loss
