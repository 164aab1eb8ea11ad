"""world_size-2 gloo test of the N>1 host logic: replica sharding and the single all-gather of
episode returns (the only collective of the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlgpuschedule_b200 import distributed as rd


def test_shard_partitions_exactly():
    for n in (1, 7, 512, 4096, 4097):
        for w in (1, 2, 3, 8):
            blocks = [rd.shard(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    first, count = rd.shard(n_total, world, rank)
    local = -(torch.arange(first, first + count, dtype=torch.int64) * 1000 + 7)   # fake per-replica returns
    full = rd.gather_returns(local, n_total)
    q.put((rank, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [8, 9])
def test_gather_returns_gloo_world2(n_total):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [-(i * 1000 + 7) for i in range(n_total)]
    assert got[0] == expect and got[1] == expect
