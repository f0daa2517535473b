"""N>1 host path on CPU: two gloo ranks shard a batch of code blocks, decode their ranges (here with the
oracle standing in for the per-rank GPU), and rank 0 gathers bytes + the max-over-ranks timing.  Checks the
partition, the gather order and the timing reduction used by bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import __graft_entry__ as ge
    import oracle_libs as ol
    pkg = ge.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, n, nit = 512, 11, 3
    bits, llr = pkg.vectors.make_blocks(n, K, 1.092, 100, seed=5)   # same batch on every rank
    counts = [pkg.sharding.shard_bounds(n, world, r)[1] - pkg.sharding.shard_bounds(n, world, r)[0] for r in range(world)]
    a, b = pkg.sharding.shard_bounds(n, world, rank)
    mine = ol.port_run_all(llr[a:b], K, nit)
    full = pkg.sharding.gather_to_rank0(dist, mine, counts)
    t = pkg.sharding.max_over_ranks(dist, 1.0 + rank)
    if rank == 0:
        want = ol.port_run_all(llr, K, nit)
        np.save(os.path.join(tmp, "ok.npy"), np.array([int(np.array_equal(full, want)), int(t == float(world))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok.tolist() == [1, 1]


def test_shard_bounds_cover_everything(pkg):
    for n in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 4, 8):
            b = [pkg.sharding.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1
    cuts = pkg.sharding.balanced_bounds([6144] * 10 + [40] * 100, 2)
    assert cuts[0] == 0 and cuts[-1] == 110 and 4 <= cuts[1] <= 7
