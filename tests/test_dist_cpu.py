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
