"""CPU (gloo, world_size 2): the multi-GPU leg of the path — contiguous batch sharding, rank-invariant per-sample
seeds, and the single all-gather of output latents (bench.py's N>1 branch)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagdressing_b200.parallel import gather_latents, sample_seeds, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, global_batch, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(global_batch, rank, world)
    seeds = sample_seeds(1234, lo, hi)
    # stand-in for the per-rank sampler: a deterministic function of the per-sample seed only
    local = torch.stack([torch.randn(4, 8, 8, generator=torch.Generator().manual_seed(s)) for s in seeds])
    full = gather_latents(local, global_batch)
    # by value (numpy pickles its bytes): a torch tensor would travel as a shared-memory handle that the parent may only
    # open after this process has already exited (seen as a sporadic FileNotFoundError in the parent)
    out_q.put((rank, lo, hi, full.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [4, 5])
def test_shard_and_gather_world2(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.randn(4, 8, 8, generator=torch.Generator().manual_seed(s))
                        for s in sample_seeds(1234, 0, global_batch)])
    spans = sorted((lo, hi) for _, lo, hi, _ in res)
    assert spans[0][0] == 0 and spans[-1][1] == global_batch and spans[0][1] == spans[1][0]
    for _, _, _, full in res:
        assert torch.equal(torch.from_numpy(full), want)  # every rank holds the same, rank-count-invariant result


def test_shard_range_covers_everything():
    for B in (1, 7, 8, 64):
        for W in (1, 2, 4, 8):
            got = []
            for r in range(W):
                lo, hi = shard_range(B, r, W)
                got += list(range(lo, hi))
            assert got == list(range(B))


# ------------------------------------------------------------------------------------------------ training: gradient all-reduce
def _train_worker(rank, world, port, out_q, accum=1):
    """Data-parallel step of FlatAdamW (imagdressing_b200/train.py; reference train.py:601-609 DeepSpeed gradient reduction):
    each rank back-propagates its own shard, the per-bucket hooks all-reduce the flat gradient buffer, the update uses the
    mean gradient — both ranks must end with the parameters of a single-process step on the whole batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulated_ops
    from imagdressing_b200 import train

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3)).to(torch.bfloat16)
    net[3].requires_grad_(False)  # a frozen tail layer: no hook fires for it, nothing is reduced for it
    opt = train.FlatAdamW(net.parameters(), lr=1e-2, weight_decay=0.0, bucket_bytes=32, step_fn=emulated_ops.adamw_step,
                          accumulation_steps=accum)
    assert opt._dist and len(opt._buckets) >= 2
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    lo, hi = shard_range(8, rank, world)
    per = (hi - lo) // accum
    reduces = []
    real = dist.all_reduce
    dist.all_reduce = lambda *a, **k: (reduces.append(opt._micro), real(*a, **k))[1]
    for _ in range(2):
        for k in range(accum):  # train.py:606: the shard in `accum` micro-batches, one update
            opt.zero_grad()
            net(x[lo + k * per:lo + (k + 1) * per]).float().square().mean().backward()
            assert opt.step() is (k == accum - 1)
    dist.all_reduce = real
    assert reduces and all(m == accum - 1 for m in reduces)  # gradients cross ranks in the window's last micro-step only
    out_q.put((rank, opt.param.float().numpy().copy(), opt.grad.float().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("accum", [1, 2])
def test_flat_adamw_gradient_allreduce_world2(accum):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulated_ops
    from imagdressing_b200 import train

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, accum)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the mean loss over the whole batch has the mean of the shard gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3)).to(torch.bfloat16)
    net[3].requires_grad_(False)
    opt = train.FlatAdamW(net.parameters(), lr=1e-2, weight_decay=0.0, bucket_bytes=32, step_fn=emulated_ops.adamw_step)
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    for _ in range(2):
        opt.zero_grad()
        net(x).float().square().mean().backward()
        opt.step()
    p0, p1 = torch.from_numpy(res[0][1]), torch.from_numpy(res[1][1])
    assert torch.equal(p0, p1)  # ranks agree bit for bit
    n = min(p0.numel(), opt.param.numel())  # (the flat buffer is padded to a multiple of 8 * world elements)
    assert float((p0[:n] - opt.param.float()[:n]).abs().max()) < 2e-2  # and match the whole-batch step to bf16 gradient rounding
    assert torch.equal(torch.from_numpy(res[0][2]), torch.from_numpy(res[1][2]))  # reduced gradient buffers identical


def _zero_worker(rank, world, port, out_q):
    """shard_states=True (ZeRO-1 style): each rank owns the fp32 master / moments of its slice only; the gathered bf16
    parameters must equal the unsharded data-parallel result."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulated_ops
    from imagdressing_b200 import train

    res = []
    for shard in (False, True):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4)).to(torch.bfloat16)
        opt = train.FlatAdamW(net.parameters(), lr=1e-2, weight_decay=0.01, bucket_bytes=32, step_fn=emulated_ops.adamw_step,
                              shard_states=shard)
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
        lo, hi = shard_range(8, rank, world)
        for _ in range(3):
            opt.zero_grad()
            net(x[lo:hi]).float().square().mean().backward()
            opt.step()
        res.append((opt.param.float().numpy().copy(), opt.master.numel(), opt.param.numel()))
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_adamw_sharded_states_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zero_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ((full, n_master_full, n_param), (sharded, n_master_shard, _)) in res:
        assert n_master_full == n_param and n_master_shard == n_param // 2  # optimizer state halved per rank
        assert (full == sharded).all()  # identical parameters, bit for bit
    assert (res[0][1][1][0] == res[1][1][1][0]).all()
