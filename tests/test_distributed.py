"""N>1 host logic on CPU: world_size-2 gloo run of the row partition + optional gather."""
import os
import sys
import socket

import numpy as np
import pytest

from cosmoprimo_b200.distributed import shard_bounds


def test_shard_bounds_cover_all_rows():
    for nrows in [0, 1, 7, 4096, 1000003]:
        for world in [1, 2, 3, 8]:
            bounds = [shard_bounds(nrows, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == nrows
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, root, queue):
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from cosmoprimo_b200.distributed import shard, gather_rows, rank_world
    from oracle import fftlog_oracle as O
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert rank_world() == (rank, world)
        n, B = 64, 7                                   # odd row count: shards of 4 and 3 rows
        k = np.geomspace(1e-3, 1e1, n)
        fun = np.exp(-k)[None, :] * (1. + np.arange(B))[:, None]
        plan = O.plan_power_to_correlation(k, ell=0)
        mine = shard(fun)
        # the per-rank transform: on a GPU box this is the cuda engine, here the oracle stands in for it
        local = O.execute(plan, mine)[1]
        full = gather_rows(local, B)
        ref = O.execute(plan, fun)[1]
        queue.put((rank, mine.shape[0], bool(np.array_equal(full, ref))))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_gather():
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, root, queue)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    res = sorted(queue.get(timeout=10) for _ in range(2))
    assert [r[1] for r in res] == [4, 3]
    assert all(r[2] for r in res)
    assert all(p.exitcode == 0 for p in procs)
