"""N>1 host logic on CPU: world_size-2 (and 4) gloo processes plan their scanline-block shards, exchange
(fake) buffer handles with all_gather_object exactly as bench.py does with CUDA-IPC handles, and every
row a shard reads must be covered by its own block or a neighbour's segment."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opticalimageprocessor_b200 import capi, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_desc(total_rows, rank, world, dY, S=30000, G=32767):
    d = capi.PanDesc()
    first, last = sharding.shard_range(total_rows, world, rank)
    d.n_ccd, d.w, d.total_rows, d.row0, d.n_rows = 3, 64, total_rows, first, last - first
    d.fold_half, d.section_rows, d.row_guard = 4, S, G
    for i in range(3):
        d.ccd[i].fmt, d.ccd[i].n_seg, d.ccd[i].shifted = capi.FMT_BE16, 1, int(i > 0)
        d.ccd[i].dX, d.ccd[i].dY = [0.0, 1.37, -0.83][i], dY[i]
    return d


def _worker(rank, world, port, total_rows, dY, S, G, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = _make_desc(total_rows, rank, world, dY, S, G)
        # "handles": what a rank would export for its 3 CCD blocks
        mine = [f"rank{rank}-ccd{i}".encode() for i in range(3)]
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        opened = []

        def peer(r, i):
            assert allh[r][i] == f"rank{r}-ccd{i}".encode()
            opened.append((r, i))
            return 0x1000000 * (r + 1) + 0x1000 * i  # stands in for the mapped pointer

        req = sharding.attach_segments(d, 3, total_rows, world, rank, [0x10 + i for i in range(3)], 128, peer)
        ok = all(sharding.covers(d, i) for i in range(3))
        # reductions the data path would need: just the timing max / counters (tiny all-reduce)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, req, sorted(set(opened)), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total_rows,S,G", [(2, 65536, 30000, 32767), (4, 131072, 30000, 32767), (2, 3000, 400, 450)])
def test_shard_planning_gloo(world, total_rows, S, G):
    dY = [0.0, -2.61, 3.19]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, dY, S, G, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    for rank, ok, req, opened, tmax in res:
        assert ok, f"rank {rank}: a needed row is not covered"
        assert tmax == float(world)
        # CCD 0 is not shifted: never needs a neighbour
        assert req[0] == []
        for i in (1, 2):
            for r in req[i]:
                assert r != rank
    # dY = -2.61 reads rows above the block: every rank but 0 needs its upper neighbour for CCD 1
    by_rank = {r[0]: r[2] for r in res}
    for rank in range(1, world):
        assert rank - 1 in by_rank[rank][1]
    # dY = +3.19 reads rows below: every rank but the last needs its lower neighbour for CCD 2
    for rank in range(world - 1):
        assert rank + 1 in by_rank[rank][2]


def test_last_rank_stale_rows_owner():
    """the partial last section's stale rows live ~30000 lines up: they may belong to a non-adjacent rank"""
    world, total = 8, 8 * 32768
    d = _make_desc(total, world - 1, world, [0.0, -2.61, 3.19])
    req = sharding.peer_requirements(d, 3, total, world, world - 1)
    (f, l), (sf, sl) = sharding.rows_needed(d, 2)
    assert sl > sf, "dY > 0 with a partial last section must read stale rows"
    owner = [r for r in range(world) if sharding.shard_range(total, world, r)[0] <= sf < sharding.shard_range(total, world, r)[1]]
    assert owner[0] in req[2] or owner[0] == world - 1
