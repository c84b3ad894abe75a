"""Whole-step CUDA graph of the Soft-IntroVAE training iteration.

One training step issues ~1150 kernel launches (13 network forwards, 2 backwards, 2 Adam steps); on small layers
the Python/driver launch path, not the GPU, sets the pace.  ``GraphedTrainStep`` captures
``trainer.soft_intro_train_step`` once -- forward, backward, loss assembly and both (capturable) Adam updates -- and
replays it per step with a single launch.  What makes the capture valid:

  * every libsivae call takes raw pointers + the current stream and never syncs or allocates;
  * torch allocates all intermediates from the graph's private pool, so replayed addresses (and the TMA tensor
    maps encoded from them at capture time) stay valid;
  * dropout masks are keyed by a device-side epoch counter advanced inside the graph (functional.begin_step);
  * reparameterisation noise is drawn inside ``sivae_reparam_draw_fwd`` from Philox keyed by the same device-side counter;
  * weight re-packing (fp32 -> bf16 tap-major) is part of the captured stream, after each optimiser step.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import functional as F
from . import trainer as T


class GraphedTrainStep:
    """Single process: ONE graph for the whole iteration.

    With ``reducer_e`` / ``reducer_d`` (``parallel.FlatGradReducer``, N > 1) the iteration is captured as THREE
    graphs sharing one memory pool, with the two gradient exchanges issued between the replays on the same stream:

        graph 0: E-phase forwards + lossE.backward()        (gradients accumulate into the flat exchange buffer)
        NCCL    : all-reduce(avg) of the encoder buffer     (one call, ~28 MB)
        graph 1: Adam(E) + repack, D-phase forwards + lossD.backward()
        NCCL    : all-reduce(avg) of the decoder buffer
        graph 2: Adam(D)

    NCCL stays outside the captured regions (capturing hook-launched collectives deadlocked on this stack), the
    host issues 5 launches per step instead of ~1500, and the stream order carries every dependency.
    """

    def __init__(self, model, optimizer_e, optimizer_d, real_example: torch.Tensor, noise_example: torch.Tensor,
                 hp: Optional[T.StepHyper] = None, warmup: int = 3, reducer_e=None, reducer_d=None):
        for opt in (optimizer_e, optimizer_d):
            if getattr(opt, "graph_safe", False):      # optim.FusedAdam: lr / step counter live on the device
                continue
            for g in opt.param_groups:
                if not g.get("capturable", False):
                    raise ValueError("GraphedTrainStep needs optim.FusedAdam or torch optimisers with capturable=True")
        self.model, self.opt_e, self.opt_d, self.hp = model, optimizer_e, optimizer_d, hp or T.StepHyper()
        self.red_e, self.red_d = reducer_e, reducer_d
        self.split = reducer_e is not None or reducer_d is not None
        # SIVAE_GRAPH_NCCL=1: capture the two all-reduces INSIDE one whole-step graph (on the update side stream, under
        # the first decoder passes of the D phase) instead of issuing them between three graph replays
        self.capture_nccl = self.split and os.environ.get("SIVAE_GRAPH_NCCL", "0") == "1"
        if self.split:
            for r in (reducer_e, reducer_d):
                if r is None or not getattr(r, "needs_persistent_grads", False):
                    raise ValueError("the multi-rank graph path needs parallel.FlatGradReducer for both phases")
        self.real = real_example.clone()
        self.noise = noise_example.clone()
        dev = self.real.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                # eager warm-up: also builds the flat gradient buffers (first finish()) and the Adam state
                T.soft_intro_train_step(model, self.real, self.noise, optimizer_e, optimizer_d, self.hp,
                                        reducer_e, reducer_d)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # With optim.FusedAdam the plain bf16 weight packs made by the warm-up stay valid for ever (the update kernel
        # rewrites them in place, torch's version counter never moves): keep them, so the captured step contains no
        # first-use pack kernels.  With torch optimisers every weight must be (re)packed inside the graph at its first
        # use after an update: drop the warm-up packs.
        self.keep_packs = all(getattr(o, "graph_safe", False) for o in (optimizer_e, optimizer_d))
        if not self.keep_packs:
            F._pack_cache.clear()
        if not self.split:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.out = T.soft_intro_train_step(model, self.real, self.noise, optimizer_e, optimizer_d, self.hp)
            self.graphs = [self.graph]
        elif self.capture_nccl:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.out = T.soft_intro_train_step(model, self.real, self.noise, optimizer_e, optimizer_d, self.hp,
                                                   reducer_e, reducer_d)
            self.graphs = [self.graph]
            self.split = False
        else:
            g0, g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0, capture_error_mode="thread_local"):
                out, z = T.soft_intro_phase_e(model, self.real, self.noise, optimizer_e, self.hp, reducer_e)
            pool = g0.pool()
            with torch.cuda.graph(g1, pool=pool, capture_error_mode="thread_local"):
                optimizer_e.step()
                out.update(T.soft_intro_phase_d(model, self.real, self.noise, z, optimizer_d, self.hp, reducer_d))
            with torch.cuda.graph(g2, pool=pool, capture_error_mode="thread_local"):
                optimizer_d.step()
            self.out = out
            self.graphs = [g0, g1, g2]
            self.graph = g0
        if not self.keep_packs:
            F._pack_cache.clear()   # the captured packs live in the graph's private pool; do not reuse them eagerly
        else:
            # the graph reads the cached pack tensors by address: pin them for the lifetime of this object, and remember
            # each weight's version so an in-place load (``load_state_dict``) between replays is noticed
            self._pinned_packs = [(v[0], key[1], v[3]) for key, v in F._pack_cache.items()]
            self._pack_versions = [(ref, ref()._version) for ref, _, _ in self._pinned_packs if ref() is not None]

    def refresh_packs(self):
        """Recompute the pinned bf16 weight packs IN PLACE from the current fp32 weights.  Needed after the weights
        were overwritten behind FusedAdam's back (``model.load_state_dict`` of a checkpoint, a broadcast): the captured
        kernels read the packs by address and only the captured Adam update rewrites them.  ``__call__`` does this
        automatically when a weight's version counter moved."""
        if not self.keep_packs:
            return
        from . import kernels as K
        with torch.no_grad():
            for ref, up, packs in self._pinned_packs:
                w = ref()
                if w is None:
                    continue
                new = K.pack_upconv3_weights(w.detach().contiguous()) if up else K.pack_conv3_weights(w.detach().contiguous())
                for dst, src in zip(packs, new):
                    dst.copy_(src)
        self._pack_versions = [(ref, ref()._version) for ref, _, _ in self._pinned_packs if ref() is not None]

    def __call__(self, real_batch: torch.Tensor, noise_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Copies the batch into the captured input buffers (device or pinned-host source) and replays.
        The returned dict holds the captured loss tensors: values are overwritten by the next call."""
        self.real.copy_(real_batch, non_blocking=True)
        self.noise.copy_(noise_batch, non_blocking=True)
        if self.keep_packs and any(ref() is not None and ref()._version != v for ref, v in self._pack_versions):
            self.refresh_packs()
        for opt in (self.opt_e, self.opt_d):
            if hasattr(opt, "sync_lr"):
                opt.sync_lr()          # LR-scheduler changes reach the device scalar the captured step reads
        if not self.split:
            self.graph.replay()
        else:
            self.graphs[0].replay()
            self.red_e.finish()
            self.graphs[1].replay()
            self.red_d.finish()
            self.graphs[2].replay()
        return self.out
