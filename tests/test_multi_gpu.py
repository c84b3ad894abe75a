"""Multi-rank equivalence on real hardware: two ranks over NCCL, one process per GPU, the headline
net through the multi-rank CUDA-graph path (three graphs + parallel.FlatGradReducer + optim.FusedAdam).

  * after the exchange every rank holds the AVERAGE of the two ranks' local gradients (each computed with replica-local
    BatchNorm statistics, exactly what nn.DataParallel replicas do -- main_DataParallel.py:609);
  * after k graph-replayed iterations the replicas' parameters are bit-identical.

Needs >= 2 GPUs: skipped on a single-GPU box (run with ``gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu``)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import sivae_b200
    from sivae_b200 import functional as F, parallel as P, trainer as T
    P.init_distributed("nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    bs = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]

    def fresh():
        torch.manual_seed(77)                                   # identical init on every rank
        net = sivae_b200.SoftIntroVAE(64, bs)
        net.apply(T.init_weights_he)
        return net.to(dev).train()

    g = torch.Generator(device=dev).manual_seed(100 + rank)     # rank-specific shard and noise
    real = torch.rand(2, 1, 16, 24, 16, device=dev, generator=g)
    noise = torch.randn(2, 1, 2, 3, 2, device=dev, generator=g)

    # (a) local gradients of one iteration, no exchange
    net = fresh()
    F.manual_seed(500 + rank)
    oe, od = torch.optim.SGD(net.encoder.parameters(), lr=0.0), torch.optim.SGD(net.decoder.parameters(), lr=0.0)
    T.soft_intro_train_step(net, real, noise, oe, od)
    names = [k for k, p in net.named_parameters() if p.grad is not None]
    local = torch.cat([p.grad.flatten() for k, p in net.named_parameters() if p.grad is not None]).clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world

    # (b) the same iteration with the exchange (eager, FlatGradReducer): every rank ends with the average
    net = fresh()
    F.manual_seed(500 + rank)
    oe, od = torch.optim.SGD(net.encoder.parameters(), lr=0.0), torch.optim.SGD(net.decoder.parameters(), lr=0.0)
    re_, rd_ = P.FlatGradReducer(net.encoder.parameters()), P.FlatGradReducer(net.decoder.parameters())
    T.soft_intro_train_step(net, real, noise, oe, od, None, re_, rd_)
    got = torch.cat([p.grad.flatten() for k, p in net.named_parameters() if p.grad is not None])
    assert [k for k, p in net.named_parameters() if p.grad is not None] == names
    err = float((got - expect).abs().max() / expect.abs().max())
    assert err < 1e-6, err
    assert float((gathered[0] - gathered[1]).abs().max()) > 0          # the shards really differ

    # (c) k iterations through the three-graph path + FusedAdam: replicas stay bit-identical
    net = fresh()
    F.manual_seed(900 + rank)
    oe, od = sivae_b200.FusedAdam(net.encoder.parameters(), lr=2e-4), sivae_b200.FusedAdam(net.decoder.parameters(), lr=2e-4)
    re_, rd_ = P.FlatGradReducer(net.encoder.parameters()), P.FlatGradReducer(net.decoder.parameters())
    step = sivae_b200.graph.GraphedTrainStep(net, oe, od, real, noise, warmup=1, reducer_e=re_, reducer_d=rd_)
    assert len(step.graphs) == 3
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    for _ in range(3):
        out = step(real, noise)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    mx, mn = flat.clone(), flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    assert torch.equal(mx, mn), "replicas diverged"
    assert float((flat - w0).abs().max()) > 0 and bool(torch.isfinite(flat).all())
    # BatchNorm running statistics are replica-local (different shards -> different buffers), as under DataParallel
    rm = net.state_dict()["encoder.blocks.1.0.block.1.running_mean"].clone()
    both = [torch.zeros_like(rm) for _ in range(world)]
    dist.all_gather(both, rm)
    assert not torch.equal(both[0], both[1])
    q.put((rank, float(out["lossE"]), err))
    dist.destroy_process_group()


def test_two_ranks_nccl_average_gradients_and_identical_replicas():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (l, e)) for r, l, e in (q.get(timeout=600) for _ in range(2)))
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert set(res) == {0, 1} and all(l == l for l, _ in res.values())
    print("two ranks: lossE per rank", {r: v[0] for r, v in res.items()}, " max rel error of the averaged gradient",
          max(v[1] for v in res.values()))
