"""One-process-per-GPU data parallelism replacing ``nn.DataParallel`` (main_DataParallel.py:609).

The training step shards by volumes only: every rank runs the whole network on
its local batch with *local* BatchNorm statistics (exactly what DataParallel replicas do), and the
only exchange is the gradient average -- encoder gradients after ``lossE.backward()``, decoder
gradients after ``lossD.backward()``.  ``GradReducer`` does that with bucketed, asynchronous
all-reduces launched from ``register_post_accumulate_grad_hook`` while backward is still producing
the remaining (earlier-layer) gradients, so NCCL traffic over NVLink overlaps the wgrad kernels.

Stock ``DistributedDataParallel`` does not fit this loop: 13 forwards precede 2 backwards, the
trainable half flips every phase, and some parameters never receive a gradient.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """Initialise torch.distributed from the torchrun environment.  -> (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "views", "pending", "work", "launched")

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        self.views = [self.flat[o: o + p.numel()].view(p.shape) for o, p in zip(self.offsets, params)]
        self.pending = set()
        self.work = None
        self.launched = False


class GradReducer:
    """Bucketed gradient averaging for one parameter group (encoder or decoder).

    Parameters are bucketed in *reverse* registration order (the order backward produces their
    gradients), ~``bucket_mb`` MiB of fp32 per bucket.  A bucket's all-reduce is launched
    asynchronously as soon as its last expected gradient has been accumulated; parameters that do not
    get a gradient in a given backward (unused projection convs) are learned on the
    first ``finish()`` and excluded from then on.  ``finish()`` launches whatever is left, waits,
    and writes the averaged values back into ``param.grad``.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 8.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params]
        self._index = {}
        self.buckets: List[_Bucket] = []
        self._expected = None  # set of params that actually receive gradients (learned on first finish)
        self._hooks = []
        cap = int(bucket_mb * (1 << 20) / 4)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._add_bucket(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._add_bucket(cur)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._arm()

    def _add_bucket(self, params):
        b = _Bucket(list(params))
        for p in params:
            self._index[p] = b
        self.buckets.append(b)

    def _arm(self):
        for b in self.buckets:
            exp = [p for p in b.params if self._expected is None or p in self._expected]
            b.pending = set(exp)
            b.work = None
            b.launched = False

    def _launch(self, b: _Bucket):
        if b.launched:
            return
        b.launched = True
        have = [(v, p.grad) for v, p in zip(b.views, b.params) if p.grad is not None]
        if len(have) != len(b.params):
            b.flat.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])     # one multi-tensor kernel
        if self.world > 1 and (have or self._expected is None):
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p):
        if self.world == 1:
            return
        b = self._index[p]
        if p in b.pending:
            b.pending.discard(p)
            if not b.pending and self._expected is not None:
                self._launch(b)

    def finish(self):
        """Call between ``loss.backward()`` and ``optimizer.step()``."""
        if self.world == 1:
            return
        if self._expected is None:
            self._expected = {p for p in self.params if p.grad is not None}
        for b in self.buckets:
            self._launch(b)
        inv = 1.0 / self.world
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
            have = [(p.grad, v) for v, p in zip(b.views, b.params) if p.grad is not None]
            if have:
                b.flat.mul_(inv)
                torch._foreach_copy_([g for g, _ in have], [v for _, v in have])
        self._arm()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class FlatGradReducer:
    """Gradient averaging with the ``.grad`` tensors of one parameter group living as views of ONE flat fp32
    buffer, so the exchange of a phase is a single NCCL all-reduce(avg) on a fixed address.

    This is the reducer of the CUDA-graph path (``graph.GraphedTrainStep`` with N > 1): backward accumulates
    straight into the flat buffer inside the captured graph, ``finish()`` runs *between* graph replays on the
    same stream, and the optimiser step of the next graph reads the averaged views.  Nothing is copied in or
    out.  The price is no overlap with backward -- the exchange is 28 MB per phase (< 0.2 ms over NVLink
    against a ~45 ms phase).

    The flat buffer is built on the first ``finish()`` from the parameters that actually received a gradient
    (the unused projection convs keep ``grad is None``, SURVEY Q1/Q2, so Adam keeps skipping them).  The
    training step must then clear gradients with ``zero_grad(set_to_none=False)`` --
    ``needs_persistent_grads`` tells the trainer.
    """

    needs_persistent_grads = True

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params]
        self.flat = None
        self.used: List[torch.nn.Parameter] = []

    def _build(self):
        self.flat = None
        self.used = [p for p in self.params if p.grad is not None]
        if not self.used:
            return
        n = sum(p.numel() for p in self.used)
        self.flat = torch.empty(n, dtype=self.used[0].grad.dtype, device=self.used[0].device)
        o = 0
        with torch.no_grad():
            for p in self.used:
                v = self.flat[o: o + p.numel()].view(p.shape)
                v.copy_(p.grad)
                p.grad = v
                o += p.numel()

    def bound(self) -> bool:
        """True while every ``.grad`` of the group still aliases its slot of the flat buffer.  ``Module.to()`` round
        trips (the trainer's per-epoch checkpoint through the CPU, utils/my_trainer.py:476-480) and
        ``zero_grad(set_to_none=True)`` re-allocate gradient storage and silently break the aliasing."""
        if self.flat is None:
            return False
        base, esz, o = self.flat.data_ptr(), self.flat.element_size(), 0
        for p in self.used:
            g = p.grad
            if g is None or g.device != self.flat.device or g.data_ptr() != base + o * esz:
                return False
            o += p.numel()
        return True

    def finish(self):
        """Call between ``loss.backward()`` and ``optimizer.step()``."""
        if not self.bound():
            # first call, or the views were lost: gather the current gradients into a fresh flat buffer and re-bind
            # (a CUDA graph captured over the old addresses is invalid anyway once the parameters have moved)
            self._build()
        if self.flat is None or self.world == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None):
    """Make parameters and buffers identical on every rank (DataParallel keeps replica 0's)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def all_reduce_mean_scalars(values: dict, group=None) -> dict:
    """Average a dict of scalar tensors across ranks with a single collective (logging only)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return values
    keys = sorted(values)
    flat = torch.stack([values[k].detach().float().reshape(()) for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    return {k: flat[i] for i, k in enumerate(keys)}
