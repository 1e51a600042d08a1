"""CPU tests of the multi-rank host logic (gloo, world_size 2): sharding covers every group
exactly once with balanced rows, and the forecast-frame gather reassembles rank order."""
import os
import socket

import numpy as np
import pytest

from time_series_spark_b200 import dist as pdist


def test_shard_bounds_cover_and_balance():
    rng = np.random.RandomState(0)
    lens = rng.randint(48, 97, 5000)
    offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    for ws in (1, 2, 3, 8):
        b = pdist.shard_bounds(offsets, ws)
        assert len(b) == ws and b[0][0] == 0 and b[-1][1] == 5000
        assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
        rows = [int(offsets[hi] - offsets[lo]) for lo, hi in b]
        assert max(rows) - min(rows) <= 2 * 96
    assert pdist.shard_bounds(np.array([0, 10]), 4) in ([(0, 1), (1, 1), (1, 1), (1, 1)], [(0, 0), (0, 0), (0, 1), (1, 1)],
                                                        [(0, 0), (0, 1), (1, 1), (1, 1)], [(0, 0), (0, 0), (0, 0), (0, 1)])
    assert pdist.shard_bounds(np.array([0]), 2) == [(0, 0), (0, 0)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    pdist.init_process_group("gloo")
    offsets = np.arange(0, 11 * 7, 7).astype(np.int64)          # 10 groups x 7 rows
    lo, hi = pdist.shard_bounds(offsets, ws)[rank]
    sid = np.arange(lo, hi, dtype=np.int32).repeat(3)
    ds = (np.arange(sid.size, dtype=np.int64) + 1000 * rank).astype("datetime64[ns]")
    yh = sid.astype(np.float64) * 0.5
    res = pdist.gather_rows([sid, ds, yh], dst=0)
    if rank == 0:
        q.put((res[0].tolist(), res[1].tolist(), res[2].tolist()))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sid, ds, yh = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sid == np.arange(10).repeat(3).tolist()
    assert yh == (np.arange(10).repeat(3) * 0.5).tolist()
    assert len(ds) == 30


def test_usable_cores_and_pool_sizing(monkeypatch):
    """Host threads = CPU affinity capped by the cgroup quota (not os.cpu_count()); each rank of a node sizes its pyarrow
    pools to its share of them (a 2-rank modeler run parsed its CSV slower than one rank before this: DESIGN 5)."""
    import pyarrow as pa
    uc = pdist.usable_cores()
    assert 1 <= uc["usable"] <= uc["affinity"] <= uc["logical"]
    if uc["cgroup_quota"] is not None:
        assert uc["usable"] <= max(1, int(uc["cgroup_quota"] + 0.5))
    before = (pa.cpu_count(), pa.io_thread_count())
    try:
        monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
        assert pdist.size_host_pools() == uc["usable"] == pa.cpu_count()
        monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
        n2 = pdist.size_host_pools()
        assert n2 == max(1, uc["usable"] // 2) == pa.cpu_count()
        assert 2 <= pa.io_thread_count() <= 8
        monkeypatch.setenv("LOCAL_WORLD_SIZE", str(4 * uc["usable"]))        # more ranks than threads: one each
        assert pdist.size_host_pools() == 1
    finally:
        pa.set_cpu_count(before[0])
        pa.set_io_thread_count(before[1])
