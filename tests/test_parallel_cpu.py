"""world_size-2 gloo test of the bucketed gradient reducer (the N>1 path of bench.py) on CPU:
two ranks with different data must end with identical, averaged gradients -- including a parameter
that never receives a gradient and a phase in which half the parameters are frozen."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), GLOO_SOCKET_IFNAME="lo")
    import sivae_b200
    from sivae_b200 import parallel as P
    r, w, _ = P.init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    enc = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4))
    unused = torch.nn.Linear(3, 3)                      # registered but never used (like the shortcut convs)
    dec = torch.nn.Linear(4, 2)
    params_e = list(enc.parameters()) + list(unused.parameters())
    red_e = P.GradReducer(params_e, bucket_mb=0.0001)   # tiny buckets -> several collectives
    red_d = P.GradReducer(dec.parameters(), bucket_mb=0.0001)
    P.broadcast_module_state(enc); P.broadcast_module_state(dec)
    out = {}
    for step in range(3):
        torch.manual_seed(100 + 10 * step + rank)       # rank-specific data
        x = torch.randn(7, 6)
        # phase E: decoder frozen
        for p in dec.parameters():
            p.requires_grad = False
        for p in enc.parameters():
            p.requires_grad = True
        for p in params_e:
            p.grad = None
        dec(enc(x)).pow(2).mean().backward()
        local = [p.grad.clone() for p in enc.parameters()]
        red_e.finish()
        gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
        expect = sum(gathered) / world
        got = torch.cat([p.grad.flatten() for p in enc.parameters()])
        assert torch.allclose(got, expect, atol=1e-6), (step, rank)
        assert all(p.grad is None for p in unused.parameters())
        # phase D: encoder frozen
        for p in dec.parameters():
            p.requires_grad = True
            p.grad = None
        for p in enc.parameters():
            p.requires_grad = False
        dec(enc(x)).pow(2).mean().backward()
        local = torch.cat([p.grad.flatten() for p in dec.parameters()])
        red_d.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        got = torch.cat([p.grad.flatten() for p in dec.parameters()])
        assert torch.allclose(got, sum(gathered) / world, atol=1e-6), (step, rank)
        out[step] = got.clone()
    scal = P.all_reduce_mean_scalars({"a": torch.tensor(float(rank)), "b": torch.tensor(2.0)})
    assert float(scal["a"]) == pytest.approx((world - 1) / 2) and float(scal["b"]) == 2.0
    q.put((rank, out[2].tolist()))        # plain lists: a tensor would travel as a shared-memory fd that dies with the child
    dist.destroy_process_group()


def _run_world(worker, world=2, attempts=3):
    """Spawn ``world`` gloo ranks; retry once on a rendezvous failure (the probed free port can be taken by
    another process between the probe and the bind)."""
    ctx = mp.get_context("spawn")
    last = None
    for _ in range(attempts):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = dict(q.get(timeout=180) for _ in range(world))
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"exit codes {[p.exitcode for p in procs]}")
        except Exception as ex:  # noqa: BLE001
            last = ex
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=10)
    raise last


def test_grad_reducer_world2_gloo():
    res = _run_world(_worker)
    assert res[0] == res[1]


def _flat_worker(rank, world, port, q):
    """FlatGradReducer (the reducer of the multi-rank CUDA-graph path): gradients live as views of one flat
    buffer, one all-reduce per phase; an unused parameter keeps grad None; replicas stay bit-identical."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), GLOO_SOCKET_IFNAME="lo")
    import sivae_b200
    from sivae_b200 import parallel as P
    from sivae_b200 import trainer as T
    P.init_distributed("gloo")
    torch.manual_seed(0)
    enc = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4))
    unused = torch.nn.Linear(3, 3)
    params = list(enc.parameters()) + list(unused.parameters())
    opt = torch.optim.Adam(params, lr=1e-2)
    red = P.FlatGradReducer(params)
    for step in range(4):
        if step == 2:
            # the trainer's per-epoch checkpoint moves the module through another device/dtype and back
            # (utils/my_trainer.py:476-480): gradient storage is re-allocated and the flat views are lost;
            # finish() must notice and re-bind, or the replicas silently stop exchanging gradients
            flat_before = red.flat
            enc.to(torch.float64)
            enc.to(torch.float32)
            assert not red.bound()
        torch.manual_seed(100 + 10 * step + rank)
        x = torch.randn(7, 6)
        T._zero_grad(opt, red)
        enc(x).pow(2).mean().backward()
        local = torch.cat([p.grad.flatten() for p in enc.parameters()]).clone()
        red.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        got = torch.cat([p.grad.flatten() for p in enc.parameters()])
        assert torch.allclose(got, sum(gathered) / world, atol=1e-6), (step, rank)
        assert all(p.grad is None for p in unused.parameters())
        # the gradients stay views of the flat exchange buffer across zero_grad / backward
        assert all(p.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr()
                   for p in enc.parameters())
        assert red.bound()
        if step == 2:
            assert red.flat is not flat_before
        opt.step()
    q.put((rank, torch.cat([p.detach().flatten() for p in enc.parameters()]).tolist()))
    dist.destroy_process_group()


def test_flat_grad_reducer_world2_gloo():
    res = _run_world(_flat_worker)
    assert res[0] == res[1]


def _fc_worker(rank, world, port, q):
    """The FC-latent variant (mymodel.py) through the trainer step on two ranks with different data: the kernels are
    replaced by their executable specification (CPU), gradients go through FlatGradReducer, Adam keeps the replicas
    identical; biases in front of a train-mode BatchNorm and the never-run block8 keep grad None on both ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), GLOO_SOCKET_IFNAME="lo")
    torch.set_num_threads(2)
    import sivae_b200
    from sivae_b200 import parallel as P
    from sivae_b200 import trainer as T
    from tests.emu import emulated_kernels
    P.init_distributed("gloo")
    torch.manual_seed(3)                                   # identical init on every rank
    net = sivae_b200.mymodel.SoftIntroVAE(4, 4, 8, 8, 6, latent_grid=(1, 1, 1))
    net.apply(T.init_weights_he)
    net.train()
    opt_e = torch.optim.Adam(net.encoder.parameters(), lr=2e-4)
    opt_d = torch.optim.Adam(net.decoder.parameters(), lr=2e-4)
    red_e, red_d = P.FlatGradReducer(net.encoder.parameters()), P.FlatGradReducer(net.decoder.parameters())
    hp = T.StepHyper(scale=sivae_b200.trainer_fc.SCALE)
    with emulated_kernels():
        for step in range(2):
            torch.manual_seed(50 + 7 * step + rank)        # rank-specific shard and noise
            real, noise = torch.rand(2, 1, 16, 16, 16), torch.randn(2, 6)
            terms = T.soft_intro_train_step(net, real, noise, opt_e, opt_d, hp, red_e, red_d)
            assert all(float(v) == float(v) for v in terms.values())
    none = sorted(k for k, p in net.named_parameters() if p.grad is None)
    assert any(".block8." in k for k in none) and "encoder.block2.0.bias" in none and "encoder.fc.bias" not in none
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    q.put((rank, (none, flat.tolist())))
    dist.destroy_process_group()


def test_fc_variant_world2_gloo():
    res = _run_world(_fc_worker)
    assert res[0][0] == res[1][0]          # same set of gradient-less parameters
    assert res[0][1] == res[1][1]          # bit-identical replicas after two E+D iterations
