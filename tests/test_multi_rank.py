"""The N>1 path of bench.py on CPU: world sharding (contiguous, disjoint ranges, one per rank) and the
metric reduction (MAX of times, SUM of counters) over torch.distributed with the gloo backend,
world_size 2.  The data path itself has no collective: worlds are independent."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world_size, port, out):
    import torch.distributed as dist

    import bench

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    w0, n = bench.shard(rank, world_size, 4096)
    vals = np.array([10.0 + rank, 20.0 - rank])
    sums = np.array([float(n), float(w0), 1.0, 2.0, 150.0, 0.0])
    v, s = bench.reduce_metrics(dist, vals, sums, "cpu")
    out[rank] = (w0, n, v.tolist(), s.tolist())
    dist.destroy_process_group()


def test_world_sharding_and_metric_reduction_gloo_ws2():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (a0, n0, v0, s0), (a1, n1, v1, s1) = out[0], out[1]
    assert (a0, n0) == (0, 4096) and (a1, n1) == (4096, 4096)          # contiguous, disjoint, weak scaling
    assert v0 == v1 == [11.0, 20.0]                                     # MAX over ranks
    assert s0 == s1 == [8192.0, 4096.0, 2.0, 4.0, 300.0, 0.0]           # SUM over ranks


def test_shard_ranges_cover_without_overlap():
    import bench

    for ws in (1, 2, 4, 8):
        seen = set()
        for r in range(ws):
            w0, n = bench.shard(r, ws, 100)
            ids = set(range(w0, w0 + n))
            assert not (ids & seen)
            seen |= ids
        assert seen == set(range(100 * ws))
