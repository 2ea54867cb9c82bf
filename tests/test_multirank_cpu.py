"""CPU, world_size 2 (gloo): the multi-rank plumbing of the session-sharded design — weight blob packed on rank 0,
broadcast once, every rank ends up with identical bytes; per-rank throughput is max-reduced the way bench.py does."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from livetalking_b200 import synth
    from livetalking_b200.w2l_pack import pack_state_dict
    if rank == 0:
        sd = synth.random_state_dict(0)
        blob = torch.from_numpy(np.frombuffer(pack_state_dict(sd), dtype=np.uint8).copy())
        n = torch.tensor([blob.numel()], dtype=torch.int64)
    else:
        blob, n = None, torch.zeros(1, dtype=torch.int64)
    dist.broadcast(n, 0)
    if rank != 0:
        blob = torch.empty(int(n.item()), dtype=torch.uint8)
    dist.broadcast(blob, 0)                                   # the design's only collective (NCCL on the GPU box)
    # session -> rank sharding and the max-over-ranks timing reduction of bench.py
    sessions = [s for s in range(6) if s % world == rank]
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, int(n.item()), int(blob[:16].sum()), int(blob.view(torch.uint8)[::4099].to(torch.int64).sum()), sessions, float(t.item())))
    dist.destroy_process_group()


@pytest.mark.timeout(280)
def test_weight_broadcast_and_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=260) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, n0, h0, c0, s0, t0), (r1, n1, h1, c1, s1, t1) = res
    assert n0 == n1 > 100_000_000 and h0 == h1 and c0 == c1          # identical blob on both ranks
    assert s0 == [0, 2, 4] and s1 == [1, 3, 5]                       # sessions shard round-robin, no overlap
    assert t0 == t1 == 11.0                                          # max over ranks
