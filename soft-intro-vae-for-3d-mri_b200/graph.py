"""Whole-step CUDA graph of the Soft-IntroVAE training iteration.

One training step issues ~1150 kernel launches (13 network forwards, 2 backwards, 2 Adam steps); on small layers
the Python/driver launch path, not the GPU, sets the pace.  ``GraphedTrainStep`` captures
``trainer.soft_intro_train_step`` once -- forward, backward, loss assembly and both (capturable) Adam updates -- and
replays it per step with a single launch.  What makes the capture valid:

  * every libsivae call takes raw pointers + the current stream and never syncs or allocates;
  * torch allocates all intermediates from the graph's private pool, so replayed addresses (and the TMA tensor
    maps encoded from them at capture time) stay valid;
  * dropout masks are keyed by a device-side epoch counter advanced inside the graph (functional.begin_step);
  * reparameterisation noise comes from torch's graph-safe Philox generator (``randn_like``);
  * weight re-packing (fp32 -> bf16 tap-major) is part of the captured stream, after each optimiser step.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import functional as F
from . import trainer as T


class GraphedTrainStep:
    def __init__(self, model, optimizer_e, optimizer_d, real_example: torch.Tensor, noise_example: torch.Tensor,
                 hp: Optional[T.StepHyper] = None, warmup: int = 3, reducer_e=None, reducer_d=None):
        for opt in (optimizer_e, optimizer_d):
            for g in opt.param_groups:
                if not g.get("capturable", False):
                    raise ValueError("GraphedTrainStep needs optimisers constructed with capturable=True")
        self.model, self.opt_e, self.opt_d, self.hp = model, optimizer_e, optimizer_d, hp or T.StepHyper()
        self.real = real_example.clone()
        self.noise = noise_example.clone()
        side = torch.cuda.Stream(device=self.real.device)
        side.wait_stream(torch.cuda.current_stream(self.real.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                T.soft_intro_train_step(model, self.real, self.noise, optimizer_e, optimizer_d, self.hp,
                                        reducer_e, reducer_d)
        torch.cuda.current_stream(self.real.device).wait_stream(side)
        torch.cuda.synchronize(self.real.device)
        # every weight must be (re)packed inside the graph at its first use: drop packs made by the warm-up
        F._pack_cache.clear()
        self.graph = torch.cuda.CUDAGraph()
        # with gradient reducers the NCCL all-reduces (launched from grad hooks) are captured too
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = T.soft_intro_train_step(model, self.real, self.noise, optimizer_e, optimizer_d, self.hp,
                                               reducer_e, reducer_d)
        F._pack_cache.clear()   # the captured packs live in the graph's private pool; do not reuse them eagerly

    def __call__(self, real_batch: torch.Tensor, noise_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Copies the batch into the captured input buffers (device or pinned-host source) and replays.
        The returned dict holds the captured loss tensors: values are overwritten by the next call."""
        self.real.copy_(real_batch, non_blocking=True)
        self.noise.copy_(noise_batch, non_blocking=True)
        self.graph.replay()
        return self.out
