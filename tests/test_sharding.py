"""CPU tier: the N>1 path (block partition + verdict gather) under gloo with world_size 2 and 3."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_cover_exactly():
    sh = importlib.import_module("recursive-stwo_b200.sharding")
    for n in (0, 1, 7, 256, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [sh.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("recursive-stwo_b200.sharding")
    lo, hi = sh.shard_range(n_total, rank, world)
    # stand-in verdicts: proof i "rejects at stage i % 9" when i % 5 == 0 (the gather is what is under test)
    ids = torch.arange(lo, hi)
    verdict = (ids % 5 == 0).to(torch.uint8)
    stage = torch.where(ids % 5 == 0, ids % 9, torch.zeros_like(ids)).to(torch.uint8)
    v, s = sh.gather_verdicts(verdict, stage, n_total)
    q.put((rank, v.numpy().tolist(), s.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 257), (3, 10), (2, 1)])
def test_gather_under_gloo(world, n_total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(n_total)
    want_v = (ids % 5 == 0).astype(np.uint8).tolist()
    want_s = np.where(ids % 5 == 0, ids % 9, 0).astype(np.uint8).tolist()
    for rank, v, s in got:
        assert v == want_v and s == want_s, rank


def _trace_worker(rank, world, port, n_total, dst, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("recursive-stwo_b200.sharding")
    lo, hi = sh.shard_range(n_total, rank, world)
    # stand-in trace columns: cell (p, c, r) = p * 1000 + c * 10 + r
    p = torch.arange(lo, hi, dtype=torch.int32).reshape(-1, 1, 1)
    local = p * 1000 + torch.arange(3, dtype=torch.int32).reshape(1, -1, 1) * 10 + torch.arange(4, dtype=torch.int32).reshape(1, 1, -1)
    out = sh.gather_trace_columns(local, n_total, dst=dst)
    q.put((rank, None if out is None else out.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,dst", [(2, 7, 0), (3, 10, None), (2, 8, 1)])
def test_gather_trace_columns_under_gloo(world, n_total, dst):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_trace_worker, args=(r, world, port, n_total, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pp = np.arange(n_total).reshape(-1, 1, 1)
    want = (pp * 1000 + np.arange(3).reshape(1, -1, 1) * 10 + np.arange(4).reshape(1, 1, -1)).tolist()
    for rank, out in got:
        if dst is None or rank == dst:
            assert out == want, rank
        else:
            assert out is None


def test_shard_range_c_entry_and_work_sharding():
    """stwo_b200_shard_range (the C entry a non-Python host calls; no device needed) = contiguous blocks differing by at most one; shard_by_work
    balances unequal units (proofs of different shapes) by cost"""
    import importlib
    sh = importlib.import_module("recursive-stwo_b200.sharding")
    for n in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            blocks = [sh.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sh.shard_range(10, 4, 4)
    work = [1] * 200 + [30] * 20 + [1] * 36          # twenty expensive units in the middle
    blocks = sh.shard_by_work(work, 4)
    assert blocks[0][0] == 0 and blocks[-1][1] == len(work) and all(blocks[r][1] == blocks[r + 1][0] for r in range(3))
    cost = [sum(work[lo:hi]) for lo, hi in blocks]
    assert max(cost) <= 1.25 * sum(work) / 4
    by_count = [sum(work[lo:hi]) for lo, hi in (sh.shard_range(len(work), r, 4) for r in range(4))]
    assert max(by_count) > 2 * max(cost) * 0.9
