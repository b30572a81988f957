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


# ------------------------------------------------------------------------------------------------ N1 on shards
def _stt_worker(rank, world, port, q):
    """each rank correlates (CPU oracle standing in for oip_stt_parameters) the sections its block holds and the four
    sums go through ONE small all-reduce -- the only collective of the path"""
    import numpy as np
    import oracle
    from test_phasecorr_cpu import _pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines, w, ov, ns, lps = 4800, 512, 200, 4, 500
        sa, sb = _pair(lines, ov, 1.37, -2.61, seed=3)
        pan1 = np.zeros((lines, w), np.uint16); pan2 = np.zeros((lines, w), np.uint16)
        pan1[:, w - ov:] = sa
        pan2[:, :ov] = sb
        lo, hi = sharding.shard_range(lines, world, rank)
        owners = sharding.stt_section_owner(lines, ns, lps, world)
        sums = np.zeros(4)
        for off, own in zip(sharding.stt_section_offsets(lines, ns, lps), owners):
            if own != rank:
                continue
            assert lo <= off and off + lps <= hi
            dx, dy, r = oracle.phase_correlate(pan1[off:off + lps, w - ov:].astype(np.float32), pan2[off:off + lps, :ov].astype(np.float32))
            if r >= 0.4:
                sums += [dx, dy, r, 1]
        t = torch.from_numpy(sums)
        dist.all_reduce(t)
        whole_rows, whole_mean = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, sections=ns, lines_per_section=lps)
        q.put((rank, owners, sharding.stt_combine(t.tolist()), whole_mean))
    finally:
        dist.destroy_process_group()


def test_stt_sections_on_two_shards_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stt_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, owners, mean, whole in res:
        assert owners == [0, 0, 1, 1]
        assert mean is not None and all(abs(a - b) < 1e-12 for a, b in zip(mean, whole))


def _stt_straddle_worker(rank, world, port, q):
    """a section that straddles two scanline blocks must still be correlated -- its rows of the two overlap
    slices are sent to the rank that holds its first line (sharding.stt_gather_straddling over gloo), so the sharded
    mean equals the whole-strip mean (ref stitcher.h:166-199 correlates every section)"""
    import numpy as np
    import oracle
    from test_phasecorr_cpu import _pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines, w, ov, ec, ns, lps = 4800, 512, 200, 6, 3, 1000
        sa, sb = _pair(lines, ov, 1.37, -2.61, seed=4)
        pan1 = np.zeros((lines, w), np.uint16); pan2 = np.zeros((lines, w), np.uint16)
        pan1[:, w - ov:] = sa
        pan2[:, :ov] = sb
        lo, hi = sharding.shard_range(lines, world, rank)
        t1, t2 = torch.from_numpy(pan1[lo:hi].copy()), torch.from_numpy(pan2[lo:hi].copy())     # the shard only
        ranges = [sharding.shard_range(lines, world, r) for r in range(world)]
        plan = sharding.stt_section_plan(lines, ns, lps, ranges)
        sums = np.zeros(4)
        n_whole = 0
        for off, owner, pieces in plan:
            if len(pieces) == 1 and owner == rank:
                n_whole += 1
                dx, dy, r = oracle.phase_correlate(pan1[off:off + lps, w - ov:w - ec].astype(np.float32), pan2[off:off + lps, ec:ov].astype(np.float32))
                if r >= 0.4:
                    sums += [dx, dy, r, 1]
        got = sharding.stt_gather_straddling(t1, t2, lo, rank, plan, (w - ov, w - ec), (ec, ov))
        for idx, a, b in got:
            dx, dy, r = oracle.phase_correlate(a.numpy().astype(np.float32), b.numpy().astype(np.float32))
            if r >= 0.4:
                sums += [dx, dy, r, 1]
        t = torch.from_numpy(sums)
        dist.all_reduce(t)
        _, whole_mean = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, edge_cols=ec, sections=ns, lines_per_section=lps)
        q.put((rank, [(o, [p[0] for p in pc]) for _, o, pc in plan], n_whole, len(got), sharding.stt_combine(t.tolist()), whole_mean, t.tolist()[3]))
    finally:
        dist.destroy_process_group()


def test_stt_straddling_section_is_gathered_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stt_straddle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, plan, n_whole, n_gathered, mean, whole, n_valid in res:
        assert plan == [(0, [0]), (0, [0, 1]), (1, [1])]          # the middle section straddles: owned by rank 0
        assert n_whole == 1 and n_gathered == (1 if rank == 0 else 0)
        assert n_valid == 3.0
        assert mean is not None and all(abs(a - b) < 1e-12 for a, b in zip(mean, whole))


def test_stt_section_plan_rejects_missing_rows():
    with pytest.raises(ValueError):
        sharding.stt_section_plan(4800, 3, 1000, [(0, 2000), (2400, 4800)])


def test_stt_section_owner_reports_straddlers():
    # reference defaults on a 4 x 65536-line strip: 16000-line sections, some cross a block boundary
    total, world = 4 * 65536, 4
    offs = sharding.stt_section_offsets(total, 10, 16000)
    own = sharding.stt_section_owner(total, 10, 16000, world)
    gap = (total - 160000) // 11
    assert offs[0] == gap and offs[1] - offs[0] == gap + 16000
    for off, o in zip(offs, own):
        inside = [r for r in range(world) if sharding.shard_range(total, world, r)[0] <= off and off + 16000 <= sharding.shard_range(total, world, r)[1]]
        assert (o == -1 and not inside) or inside == [o]
    assert sharding.stt_combine([0, 0, 0, 0]) is None and sharding.stt_combine([2.0, 4.0, 1.0, 2.0]) == (1.0, 2.0, 0.5)


def test_row_ranges_are_tight_and_inside_the_bounding_ranges():
    """oip_pan_row_ranges (what the host-buffer pipeline copies per row block): disjoint sorted ranges inside the bounding
    ranges of oip_pan_rows_needed, covering their end points, and far smaller than the bounding stale range of the
    partial last section (C2: one stale row ~27000 rows above the block instead of the whole previous section)"""
    import ctypes as C
    L = capi.load()
    total = 32768
    for r0, nr in [(0, 4096), (12288, 4096), (28672, 4096), (30000, 2768), (0, total)]:
        d = _make_desc(total, 0, 1, [0.0, -2.61, 3.19])
        d.row0, d.n_rows = r0, nr
        for i in range(3):
            (f, l), (sf, sl) = sharding.rows_needed(d, i)
            rg = (C.c_int64 * 8)()
            n = C.c_int()
            capi.check(L.oip_pan_row_ranges(C.byref(d), i, rg, 4, C.byref(n)))
            rs = [(rg[2 * k], rg[2 * k + 1]) for k in range(n.value)]
            assert rs == sorted(rs) and all(a < b for a, b in rs)
            assert all(b0 < a1 for (_, b0), (a1, _) in zip(rs, rs[1:]))
            lo, hi = min(f if l > f else 1 << 60, sf if sl > sf else 1 << 60), max(l, sl)
            assert rs[0][0] == lo and rs[-1][1] == hi
            for a, b in rs:
                assert (f <= a and b <= l) or (sf <= a and b <= sl) or (min(f, sf) <= a and b <= max(l, sl))
            if (r0, nr, i) == (28672, 4096, 2):
                assert sl - sf > 20000 and sum(b - a for a, b in rs) < nr + 16


def test_mss_sections_follow_the_reference_loop_and_partition_over_ranks():
    import numpy as np
    import oracle
    for lines, lps, ov, off, keep in [(65536, 20000, 520, 0, False), (65536, 20000, 520, 100, True), (5000, 1800, 200, 0, False),
                                      (41000, 20000, 520, 0, False)]:
        secs = sharding.mss_sections(lines, lps, ov, off, keep, 1500)
        # output rows are contiguous and add up to the reference's processedLines
        assert [s[3] for s in secs] == list(np.cumsum([0] + [s[4] for s in secs[:-1]]))
        assert all(s[2] == ov for s in secs[1:]) and secs[0][2] == (0 if keep else ov)
        assert all(b[0] - a[0] == lps - ov for a, b in zip(secs, secs[1:]))
        for world in (1, 2, 3, 8):
            parts = [sharding.mss_rank_sections(secs, world, r) for r in range(world)]
            assert sorted(s for p in parts for s in p) == sorted(secs)
            for p in parts:  # contiguous runs
                idx = [secs.index(s) for s in p]
                assert idx == list(range(idx[0], idx[0] + len(idx))) if idx else True
    # against the oracle's own section loop on a small strip: per-section calls == whole call
    rng = np.random.default_rng(0)
    lines, wb = 700, 64
    planes = [rng.integers(0, 4096, (lines, wb), dtype=np.uint16) for _ in range(4)]
    cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
    cY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8] for b in range(4)]
    n, whole = oracle.band_align(planes, cX, cY, lines_per_section=300, overlap=40, min_process_lines=100)
    secs = sharding.mss_sections(lines, 300, 40, 0, False, 100)
    assert sum(s[4] for s in secs) == n
    for (o, m, y0, o0, no) in secs:
        k, part = oracle.band_align([p[o:o + m] for p in planes], cX, cY, lines_per_section=300, overlap=40, keep_leading=(y0 == 0),
                                    min_process_lines=41)
        assert k == no and np.array_equal(part[:k], whole[o0:o0 + no])


# ------------------------------------------------------------------------------------------------ stage 1 on byte-range shards
def _stage1_file(seed=3, prefix=b"", restart_at=None):
    import numpy as np
    from opticalimageprocessor_b200 import synth
    imdt, _ = synth.make_imdt(6, 32, 8, seed=seed)
    imtr = synth.imtr_frames(imdt, chid=0x22)
    if restart_at is not None:                      # a frame with sequence number 0: the frame after it re-creates the IMDT file
        imtr[restart_at, 4:8] = 0
        synth.refresh_imtr_crc(imtr)
    imtr[40, 300] ^= 0x10                            # one IMTR frame with a bad CRC (its AOS frames are fine)
    aos = synth.aos_frames(imtr.reshape(-1))
    return synth.build_aos_file(aos, empty_every=7, bad_crc_at={3, 50}, bad_inject_at={9}, prefix=prefix)


def _oracle_imtr_shard(buf, payload_off, skip, n_frames):
    """CPU stand-in for oip_imtr_deframe_shard(prev_seq unknown): the oracle's re-framing on the shard's own stream"""
    import numpy as np
    import oracle
    stream = np.concatenate([buf[int(o):int(o) + 880] for o in payload_off]) if len(payload_off) else np.zeros(0, np.uint8)
    stream = stream[skip:skip + 882 * n_frames]
    pad = (-stream.size) % 880
    stream = np.concatenate([stream, np.zeros(pad, np.uint8)])
    offs = np.arange(0, stream.size, 880, dtype=np.uint64)
    imdt, st = oracle.imtr_deframe(stream, offs)
    assert st[0] == n_frames
    # the oracle applied "previous seq = 0" to the first valid frame: take that rule out again, keep the sequence numbers
    seqs = []
    for q in range(n_frames):
        fr = stream[882 * q:882 * q + 882]
        ok = fr[:4].tobytes() == b"\x49\x54\xCE\x1F" and fr[878:882].tobytes() == b"\x2E\xE9\xC8\xFD" and fr[9] == 0x22 and \
            oracle.crc16(fr[:876]) == (int(fr[876]) << 8 | int(fr[877]))
        if ok:
            seqs.append(int.from_bytes(fr[4:8].tobytes(), "big"))
    local_restart = max([k for k in range(1, len(seqs)) if seqs[k - 1] == 0], default=-1)
    info = dict(n_frames=n_frames, n_valid=len(seqs), bad=[int(st[2]), int(st[3]), int(st[4]), int(st[5])],
                first_seq=seqs[0] if seqs else -1, last_seq=seqs[-1] if seqs else -1,
                gaps=sum(1 for k in range(1, len(seqs)) if seqs[k - 1] + 1 != seqs[k]),
                restarts=sum(1 for k in range(1, len(seqs)) if seqs[k - 1] == 0), local_restart=local_restart,
                first_chid=int(st[7]), imdt_bytes=int(imdt.size))
    return imdt, info


def _run_stage1_sharded(buf, world, ranges, aos_shard, imtr_shard):
    """sequential simulation of `world` ranks with the two exchanges of SURVEY 8e spelled out"""
    import numpy as np
    n = buf.size
    subs = [buf[a:min(n, b + 1023)] for a, b in ranges]
    # round 0: every shard from its first byte; "all-gather" of (carry_in, carry_out, n_valid); re-run where the assumption was wrong
    res = [aos_shard(subs[r], ranges[r][1] - ranges[r][0], 0) for r in range(world)]
    carry_in = [0] * world
    for _ in range(world):
        want = [0] + [res[r][2] for r in range(world - 1)]
        if want == carry_in:
            break
        for r in range(world):
            if want[r] != carry_in[r]:
                carry_in[r] = want[r]
                res[r] = aos_shard(subs[r], ranges[r][1] - ranges[r][0], carry_in[r])
    else:
        raise AssertionError("carries did not settle")
    counters = sum(r_[1] for r_ in res)
    n_valid = [int(r_[1][0]) for r_ in res]
    # halo payloads: the first two payloads of every rank travel ("all-gather" of 1760 bytes + n_valid)
    heads = [[subs[r][int(o):int(o) + 880] for o in res[r][0][:2]] for r in range(world)]
    pieces, infos = [], []
    for r in range(world):
        f0, nf, skip, halo = sharding.imtr_shard_frames(n_valid, r)
        following = [h for q in range(r + 1, world) for h in heads[q]][:halo]
        assert len(following) == halo
        ext = np.concatenate([subs[r]] + following) if following else subs[r]
        offs = list(res[r][0]) + [subs[r].size + 880 * k for k in range(halo)]
        piece, info = imtr_shard(ext, np.array(offs, np.uint64), skip, nf)
        pieces.append(piece)
        infos.append(info)
    keep, stats = sharding.imtr_combine(infos)
    all_off = np.concatenate([res[r][0].astype(np.uint64) + np.uint64(ranges[r][0]) for r in range(world)])
    imdt = np.concatenate([pieces[r] for r in range(world) if keep[r]] or [np.zeros(0, np.uint8)])
    return all_off, counters, imdt, stats, carry_in


@pytest.mark.parametrize("prefix,restart_at", [(b"", None), (b"\x00" * 13, None), (b"", 77), (b"\x07" * 5, 150)])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_stage1_byte_range_shards_equal_the_whole_file(world, prefix, restart_at):
    """AOS scan + IMTR re-framing on byte-range shards (carry resolution, payload-count prefix, seq rules across ranks)
    == the sequential whole-file oracle: payload list, the 3 counters, IMDT bytes and the 9 IMTR stats"""
    import numpy as np
    import oracle
    buf = _stage1_file(prefix=prefix, restart_at=restart_at)
    off_w, cnt_w = oracle.aos_scan(buf)
    imdt_w, st_w = oracle.imtr_deframe(buf, off_w)

    def aos_shard(sub, own, carry):
        o, c, nxt = oracle.aos_scan_range(sub, carry, own)
        return o, c, max(0, nxt - own)
    for ranges in (sharding.aos_shard_ranges(buf.size, world, len(prefix)),                       # on the frame cadence: no carries
                   [(buf.size * r // world + (3 if r else 0), buf.size * (r + 1) // world + (3 if r + 1 < world else 0)) for r in range(world)]):
        off, cnt, imdt, stats, carries = _run_stage1_sharded(buf, world, ranges, aos_shard, _oracle_imtr_shard)
        assert np.array_equal(off, off_w) and cnt.tolist() == cnt_w.tolist()
        assert stats == st_w.tolist(), (stats, st_w.tolist())
        assert np.array_equal(imdt, imdt_w)
    assert any(c for c in carries)    # the second partition cuts through frames


def _carry_worker(rank, world, port, q):
    """aos_resolve_carries over gloo: ONE all-gather per round, the rank with a wrong assumption scans again"""
    import numpy as np
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        buf = _stage1_file()
        a, b = [(0, 70001), (70001, buf.size)][rank]
        sub = buf[a:min(buf.size, b + 1023)]
        calls = []

        def scan(carry):
            calls.append(carry)
            o, c, nxt = oracle.aos_scan_range(sub, carry, b - a)
            return (o + np.uint64(a), c), max(0, nxt - (b - a)), int(c[0])
        (off, cnt), carry_in, n_valid_all = sharding.aos_resolve_carries(scan, world, rank)
        t = torch.from_numpy(cnt.copy())
        dist.all_reduce(t)                                  # the 3-counter all-reduce
        gathered = [None] * world
        dist.all_gather_object(gathered, off.tolist())
        q.put((rank, calls, carry_in, n_valid_all, t.tolist(), sum(gathered, [])))
    finally:
        dist.destroy_process_group()


def test_aos_carry_resolution_gloo():
    import numpy as np
    import oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_carry_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    buf = _stage1_file()
    off_w, cnt_w = oracle.aos_scan(buf)
    for rank, calls, carry_in, n_valid_all, cnt, offs in res:
        assert cnt == cnt_w.tolist() and offs == off_w.tolist() and sum(n_valid_all) == cnt_w[0]
        assert calls == ([0] if rank == 0 else [0, carry_in]) and (rank == 0 or carry_in > 0)


# ------------------------------------------------------------------------------------------------
# frame index: host-only chain (oip_image_frames_chain) and the index over pieces of an IMDT stream
# ------------------------------------------------------------------------------------------------
def _np_find_hits(buf):
    """numpy stand-in for the device search (oip_image_frames_hits): ascending signature offsets + the 172 bytes there"""
    import numpy as np
    from opticalimageprocessor_b200 import synth
    sig = np.frombuffer(synth.IMG_SIG, np.uint8)
    if buf.size < 4:
        return np.zeros(0, np.uint64), np.zeros((0, 172), np.uint8)
    m = (buf[:-3] == sig[0]) & (buf[1:-2] == sig[1]) & (buf[2:-1] == sig[2]) & (buf[3:] == sig[3])
    off = np.flatnonzero(m).astype(np.uint64)
    pad = np.concatenate([buf, np.zeros(172, np.uint8)])
    tr = np.stack([pad[int(o):int(o) + 172] for o in off]) if off.size else np.zeros((0, 172), np.uint8)
    return off, tr


def _np_unpack(imdt, ents, n, tc, tl):
    """aux / PAN / MSS rasters from a frame table (what oip_unpack_frames does), to compare a table with the oracle's output"""
    import numpy as np
    W = 8 * tc
    aux = np.zeros((n, 192 * tl), np.uint8)
    pan = np.zeros((n * 4 * tl, W), np.uint16)
    mss = np.zeros((n * tl, W), np.uint16)
    for f in range(n):
        e = ents[f]
        if e.frame_off < 0:
            continue
        aux[f] = imdt[e.frame_off:e.frame_off + 192 * tl]
        for k in range(40):
            o = e.tile_off[k]
            t = imdt[o:o + tc * tl * 2].view(">u2").reshape(tl, tc)
            r, c = divmod(k, 8)
            if r < 4:
                pan[f * 4 * tl + r * tl:f * 4 * tl + (r + 1) * tl, c * tc:(c + 1) * tc] = t
            else:
                mss[f * tl:(f + 1) * tl, c * tc:(c + 1) * tc] = t
    return aux, pan, mss


def _imdt_cases():
    import numpy as np
    from opticalimageprocessor_b200 import synth
    tc, tl = 16, 4
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    a, _ = synth.make_imdt(5, tc, tl, seed=10)
    b, _ = synth.make_imdt(6, tc, tl, seed=11, skip_seqs={3}, junk_prefix=333)
    c, _ = synth.make_imdt(4, tc, tl, seed=9)
    c1 = c[100:].copy()                                   # first frame incomplete
    c2 = c.copy()
    pos = frame_bytes + 192 * tl + 64                     # a false signature inside frame 2
    c2[pos:pos + 4] = np.frombuffer(synth.IMG_SIG, np.uint8)
    c3 = c[:-50].copy()                                   # last trailer cut
    return tc, tl, frame_bytes, [a, b, c1, c2, c3]


def test_frames_chain_on_the_host_equals_the_oracle():
    """oip_image_frames_chain needs no GPU: fed with the signature table of a numpy search it reproduces the oracle's frame
    walk (stats, and the rasters its table unpacks to) -- junk prefix, sequence gap, incomplete first frame, false
    signature, cut last trailer"""
    import numpy as np
    import oracle
    from opticalimageprocessor_b200 import ops
    tc, tl, _, cases = _imdt_cases()
    for i, buf in enumerate(cases):
        n_w, aux_w, pan_w, mss_w, st_w = oracle.image_frames(buf, tc, tl)
        off, tr = _np_find_hits(buf)
        ents, st = ops.image_frames_chain(off, tr, buf.size, tc, tl)
        assert st.tolist() == st_w.tolist(), i
        aux, pan, mss = _np_unpack(buf, ents, int(st[1]), tc, tl)
        assert np.array_equal(aux, aux_w) and np.array_equal(pan, pan_w) and np.array_equal(mss, mss_w), i
    # degenerate tables
    ents, st = ops.image_frames_chain(np.zeros(0, np.uint64), np.zeros((0, 172), np.uint8), 10 ** 6, tc, tl)
    assert st.tolist() == [0, 0, 0, 0]
    ents, st = ops.image_frames_chain(np.zeros(0, np.uint64), np.zeros((0, 172), np.uint8), 100, tc, tl)
    assert st.tolist() == [0, 0, 0, 0]


def _piece_cuts(n, frame_bytes, world, variant):
    if variant == 0:                                       # equal pieces
        return [n * r // world for r in range(world + 1)]
    if variant == 1:                                       # cuts inside signatures / trailers: 1, 2, 3 bytes and 100 bytes into a trailer
        c = [0] + [min(n, (r * (n // frame_bytes) // world) * frame_bytes + frame_bytes - 172 + (1, 2, 3, 100)[r % 4]) for r in range(1, world)] + [n]
        return sorted(c)
    if variant == 2:
        return [0] + [min(n, 7 * r) for r in range(1, world)] + [n]   # tiny leading pieces (shorter than the halo)
    # empty pieces (ranks whose IMDT piece was dropped by a sequence restart): first and, for world > 2, one in the middle
    c = [0, 0] + [n * r // world for r in range(2, world)] + [n]
    if world > 2:
        c[2] = c[3]
    return c


@pytest.mark.parametrize("world", [2, 3, 5])
def test_frames_index_on_pieces_equals_the_whole_stream(world):
    """the IMDT stream cut into `world` pieces at arbitrary bytes (through signatures, through trailers, pieces shorter than
    the halo): piece + halo search, contributions in rank order, host chain == the chain over the whole stream"""
    import numpy as np
    from opticalimageprocessor_b200 import ops
    tc, tl, frame_bytes, cases = _imdt_cases()
    for i, buf in enumerate(cases):
        ents_w, st_w = ops.image_frames_chain(*_np_find_hits(buf), buf.size, tc, tl)
        for variant in range(4):
            cuts = _piece_cuts(buf.size, frame_bytes, world, variant)
            assert len(cuts) == world + 1 and cuts == sorted(cuts)
            pieces = [buf[cuts[r]:cuts[r + 1]] for r in range(world)]
            sizes = [p.size for p in pieces]
            heads = [p[:sharding.FRAME_HALO] for p in pieces]       # "all-gather" of the piece heads
            payloads = []
            for r in range(world):
                ext = np.concatenate([pieces[r], sharding.frames_piece_halo(heads, r)])
                assert ext.size <= pieces[r].size + sharding.FRAME_HALO
                payloads.append(sharding.frames_local_hits(lambda: _np_find_hits(ext), sizes, r))
            ents, st = sharding.frames_chain_all(payloads, sizes, tc, tl)
            assert st.tolist() == st_w.tolist(), (i, variant)
            for f in range(int(st[1])):
                assert ents[f].frame_off == ents_w[f].frame_off and list(ents[f].tile_off) == list(ents_w[f].tile_off) and ents[f].seq == ents_w[f].seq


def _frames_worker(rank, world, port, q):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tc, tl, frame_bytes, cases = _imdt_cases()
        buf = cases[1]
        cut = 2 * frame_bytes + 333 + frame_bytes - 172 + 2            # two bytes into the third trailer's signature
        piece = buf[:cut] if rank == 0 else buf[cut:]
        sizes = [cut, buf.size - cut]
        heads = [None] * world
        dist.all_gather_object(heads, piece[:sharding.FRAME_HALO].tobytes())
        ext = np.concatenate([piece, sharding.frames_piece_halo([np.frombuffer(h, np.uint8) for h in heads], rank)])
        ents, st = sharding.frames_index_shards(lambda: _np_find_hits(ext), sizes, rank, tc, tl)
        q.put((rank, st.tolist(), [(ents[f].frame_off, ents[f].seq, ents[f].tile_off[39]) for f in range(int(st[1]))]))
    finally:
        dist.destroy_process_group()


def test_frames_index_shards_gloo():
    """two processes, one all_gather_object of the piece heads and one of the (offset, trailer) contributions: both ranks
    end up with the whole-stream frame table"""
    import oracle
    from opticalimageprocessor_b200 import ops
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frames_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tc, tl, _, cases = _imdt_cases()
    buf = cases[1]
    ents_w, st_w = ops.image_frames_chain(*_np_find_hits(buf), buf.size, tc, tl)
    assert st_w.tolist() == oracle.image_frames(buf, tc, tl)[4].tolist()
    want = [(ents_w[f].frame_off, ents_w[f].seq, ents_w[f].tile_off[39]) for f in range(int(st_w[1]))]
    for rank, st, table in res:
        assert st == st_w.tolist() and table == want


def test_frames_chain_error_behaviour():
    """the host chain refuses what the index refuses: a JPEG-2000 compressed frame (z_ratio != 0, ref aux_separator.h:639 --
    OIP_E_UNSUPPORTED), a table that does not fit the caller's capacity (OIP_E_INVALID with the needed count
    in stats[1]), a bad geometry; a sub-image that would leave the buffer is OIP_E_RANGE"""
    import ctypes as C
    import numpy as np
    from opticalimageprocessor_b200 import ops
    from opticalimageprocessor_b200.capi import FrameEntry, FrameGeom, OipError
    tc, tl, frame_bytes, cases = _imdt_cases()
    buf = cases[0]
    off, tr = _np_find_hits(buf)
    z = tr.copy()
    z[1, 4] |= 0x05
    with pytest.raises(OipError, match="JPEG-2000") as ei:
        ops.image_frames_chain(off, z, buf.size, tc, tl)
    assert ei.value.code == capi.OIP_E_UNSUPPORTED
    lib = capi.load()
    st = (C.c_int64 * 4)()
    ents = (FrameEntry * 2)()
    g = FrameGeom(tc, tl)
    rc = lib.oip_image_frames_chain(off.ctypes.data, tr.ctypes.data, len(off), buf.size, C.byref(g), ents, 2, st)
    assert rc == capi.OIP_E_INVALID and st[1] == 5 and ents[1].frame_off == frame_bytes       # the first two entries are filled
    rc = lib.oip_image_frames_chain(off.ctypes.data, tr.ctypes.data, len(off), buf.size, C.byref(g), None, 0, st)
    assert rc == capi.OIP_OK and st[1] == 5                                                   # count-only call
    bad = FrameGeom(0, tl)
    assert lib.oip_image_frames_chain(off.ctypes.data, tr.ctypes.data, len(off), buf.size, C.byref(bad), None, 0, st) == capi.OIP_E_INVALID
    # the same table read with four times the tile width: the last sub-image of the last frame would leave the buffer
    with pytest.raises(OipError, match="leaves the buffer") as ei:
        ops.image_frames_chain(off, tr, buf.size, 4 * tc, tl)
    assert ei.value.code == capi.OIP_E_RANGE


def test_frames_chain_fuzz_against_the_oracle():
    """1000 damaged IMDT streams (ranges dropped, false signatures, cut tails, junk inserted, sequence gaps, junk prefixes):
    the host chain fed by a numpy signature search gives the oracle's frame walk -- same counters, same rasters, and an error
    exactly where the oracle refuses the stream (a sub-image that would leave the buffer)"""
    import numpy as np
    import oracle
    from opticalimageprocessor_b200 import ops, synth
    from opticalimageprocessor_b200.capi import OipError
    rng = np.random.default_rng(20221019)
    sig = np.frombuffer(synth.IMG_SIG, np.uint8)
    n_err = 0
    for it in range(1000):
        tc, tl = int(rng.choice([8, 16, 24])), int(rng.choice([1, 2, 4]))
        nfr = int(rng.integers(1, 7))
        skip = set(int(x) for x in rng.choice(np.arange(1, nfr + 2), size=int(rng.integers(0, 2)), replace=False)) if nfr > 1 else set()
        imdt, _ = synth.make_imdt(nfr, tc, tl, seed=int(rng.integers(1 << 30)), skip_seqs=skip, junk_prefix=int(rng.choice([0, 0, 5, 333])))
        buf = imdt.copy()
        for _ in range(int(rng.integers(0, 4))):
            kind = int(rng.integers(0, 4))
            if kind == 0 and buf.size > 600:
                a = int(rng.integers(0, buf.size - 500))
                buf = np.concatenate([buf[:a], buf[a + int(rng.integers(1, 500)):]])
            elif kind == 1:
                p = int(rng.integers(0, buf.size - 4))
                buf[p:p + 4] = sig
            elif kind == 2:
                buf = buf[:max(10, buf.size - int(rng.integers(1, 400)))]
            else:
                p = int(rng.integers(0, buf.size))
                buf = np.concatenate([buf[:p], rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8), buf[p:]])
        buf = np.ascontiguousarray(buf)
        try:
            n_w, aux_w, pan_w, mss_w, st_w = oracle.image_frames(buf, tc, tl)
        except AssertionError:        # the oracle's fill pass refused what its counting pass had accepted
            n_w = -1
        off, tr = _np_find_hits(buf)
        try:
            ents, st = ops.image_frames_chain(off, tr, buf.size, tc, tl)
            failed = False
        except OipError:
            failed = True
        if n_w < 0 or failed:
            assert n_w < 0 and failed, it
            n_err += 1
            continue
        assert st.tolist() == st_w.tolist(), it
        aux, pan, mss = _np_unpack(buf, ents, int(st[1]), tc, tl)
        assert np.array_equal(aux, aux_w) and np.array_equal(pan, pan_w) and np.array_equal(mss, mss_w), it
    assert 20 < n_err < 500


def test_stage1_shards_fuzz_against_the_whole_file():
    """300 random downlinks (damaged IMTR frames, sequence restarts, empty / bad-CRC / bad-inject AOS frames, byte slips,
    prefixes) cut into 2..6 shards at random bytes (shards shorter than a frame included): carry resolution, the payload-count
    prefix of the IMTR cadence and the sequence rules across ranks reproduce the sequential whole-file result -- payload
    list, the 3 counters, IMDT bytes, the 9 IMTR stats"""
    import numpy as np
    import oracle
    from opticalimageprocessor_b200 import synth
    rng = np.random.default_rng(4711)

    def aos_shard(sub, own, carry):
        o, c, nxt = oracle.aos_scan_range(sub, carry, own)
        return o, c, max(0, nxt - own)
    done = 0
    for it in range(300):
        imdt, _ = synth.make_imdt(int(rng.integers(1, 4)), 16, 4, seed=int(rng.integers(1 << 30)))
        imtr = synth.imtr_frames(imdt, chid=0x22)
        nfr = imtr.shape[0]
        for _ in range(int(rng.integers(0, 3))):
            imtr[int(rng.integers(0, nfr)), int(rng.integers(0, 882))] ^= 0x20
        if rng.random() < 0.3 and nfr > 5:
            imtr[int(rng.integers(1, nfr - 1)), 4:8] = 0
            synth.refresh_imtr_crc(imtr)
        aos = synth.aos_frames(imtr.reshape(-1))
        na = aos.shape[0]
        buf = synth.build_aos_file(aos, empty_every=int(rng.integers(3, 12)), bad_crc_at=set(int(x) for x in rng.integers(0, na, 2)),
                                   bad_inject_at=set(int(x) for x in rng.integers(0, na, 1)), prefix=bytes(int(rng.integers(0, 20)))).copy()
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, buf.size))
            if rng.random() < 0.5:
                buf = np.concatenate([buf[:p], rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8), buf[p:]])
            else:
                buf = np.concatenate([buf[:p], buf[p + int(rng.integers(1, 9)):]])
        buf = np.ascontiguousarray(buf)
        off_w, cnt_w = oracle.aos_scan(buf)
        imdt_w, st_w = oracle.imtr_deframe(buf, off_w)
        world = int(rng.integers(2, 7))
        cuts = sorted(int(x) for x in rng.integers(1, buf.size - 1, world - 1))
        if len(set(cuts)) < world - 1:
            continue
        ranges = list(zip([0] + cuts, cuts + [buf.size]))
        off, cnt, imdt_s, stats, _ = _run_stage1_sharded(buf, world, ranges, aos_shard, _oracle_imtr_shard)
        assert np.array_equal(off, off_w) and cnt.tolist() == cnt_w.tolist(), (it, ranges)
        assert stats == st_w.tolist() and np.array_equal(imdt_s, imdt_w), (it, ranges, stats, st_w.tolist())
        done += 1
    assert done > 250


def test_row_ranges_are_sufficient_and_tight_fuzz():
    """oip_pan_row_ranges against the oracle on 600 random strips / row shards (2-3 CCDs, random folds, section and guard
    sizes incl. single-section strips, shifts of both signs): the oracle computes the shard's output once from the whole
    strip and once from a strip in which every row OUTSIDE the declared ranges is random garbage -- identical, so the ranges
    hold every row the arithmetic reads (halo rows, the stale rows of a partial last section); with one row taken off
    either end of every range the two differ, so the ranges are tight"""
    import ctypes as C
    import numpy as np
    import oracle
    from opticalimageprocessor_b200 import synth
    L = capi.load()
    rng = np.random.default_rng(99)
    w = 64
    teeth = 0
    for it in range(600):
        n = int(rng.integers(2, 4))
        f = int(rng.integers(0, 9))
        G = int(rng.integers(40, 300))
        S = int(rng.integers(8, G + 1))
        total = int(rng.integers(G + 1, 1200)) if rng.random() < 0.85 else int(rng.integers(20, G + 1))
        dX = [0.0] + [float(np.round(rng.uniform(-5, 5), 2)) for _ in range(n - 1)]
        dY = [0.0] + [float(np.round(rng.uniform(-5, 5), int(rng.integers(0, 3)))) for _ in range(n - 1)]
        row0 = int(rng.integers(0, total))
        n_rows = int(rng.integers(1, total - row0 + 1))
        d = capi.PanDesc()
        d.n_ccd, d.w, d.total_rows, d.row0, d.n_rows = n, w, total, row0, n_rows
        d.fold_half, d.section_rows, d.row_guard = f, S, G
        for i in range(n):
            d.ccd[i].fmt, d.ccd[i].n_seg, d.ccd[i].shifted = capi.FMT_LE16, 1, int(i > 0)
            d.ccd[i].dX, d.ccd[i].dY = dX[i], dY[i]
        data = [rng.integers(0, 65536, (total, w), dtype=np.uint16) for _ in range(n)]
        kbs = [synth.rrc_coeffs(w, 3 + i) for i in range(n)]
        ranges = []
        for i in range(n):
            rg = (C.c_int64 * 16)()
            cnt = C.c_int()
            capi.check(L.oip_pan_row_ranges(C.byref(d), i, rg, 8, C.byref(cnt)))
            ranges.append([(rg[2 * k], rg[2 * k + 1]) for k in range(cnt.value)])

        def poisoned(rgs):
            def gen(i, a, b):
                out = rng.integers(0, 65536, (b - a, w), dtype=np.uint16)
                for ra, rb in rgs[i]:
                    lo, hi = max(a, ra), min(b, rb)
                    if hi > lo:
                        out[lo - a:hi - a] = data[i][lo:hi]
                return out
            return gen
        rows = np.arange(row0, row0 + n_rows)
        what = dict(it=it, n=n, f=f, S=S, G=G, total=total, dX=dX, dY=dY, row0=row0, n_rows=n_rows, ranges=ranges)
        want = oracle.pan_rows(lambda i, a, b: data[i][a:b], n, w, kbs, dX, dY, f, total, rows, S, G)
        assert np.array_equal(oracle.pan_rows(poisoned(ranges), n, w, kbs, dX, dY, f, total, rows, S, G), want), what
        if it % 10 == 0:      # the test has teeth: a row short at either end changes the output
            for cut in ([[(a, b - 1) for a, b in r] for r in ranges], [[(a + 1, b) for a, b in r] for r in ranges]):
                teeth += not np.array_equal(oracle.pan_rows(poisoned(cut), n, w, kbs, dX, dY, f, total, rows, S, G), want)
    assert teeth >= 110
