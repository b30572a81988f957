"""GPU parity, stage 1 (frame handling): CUDA path vs the CPU oracle.  Bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from opticalimageprocessor_b200 import synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _downlink(seed=3, n_frames=3, tc=16, tl=4, skip=()):
    imdt, truth = synth.make_imdt(n_frames, tc, tl, seed=seed, skip_seqs=skip)
    imtr = synth.imtr_frames(imdt, chid=0x22)
    aos = synth.aos_frames(imtr.reshape(-1))
    return imdt, truth, imtr, aos


@pytest.mark.parametrize("length", [0, 1, 5, 31, 32, 33, 876, 890, 1000])
def test_crc16_batch(ctx, oracle_mod, length):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(length)
    buf = rng.integers(0, 256, 8192, dtype=np.uint8)
    off = rng.integers(0, 8192 - max(length, 1), 77).astype(np.int64)
    got = ops.crc16_batch(ctx, _dev(buf), _dev(off), length).cpu().numpy()
    want = np.array([oracle_mod.crc16(buf[o:o + length]) for o in off], np.uint16)
    assert np.array_equal(got, want)


def _check_aos(ctx, oracle_mod, buf):
    from opticalimageprocessor_b200 import ops
    off_w, cnt_w = oracle_mod.aos_scan(buf)
    for fused in (1, 0):   # the single-pass cadence kernel (default) and the exhaustive search kernels it falls back to
        ctx.set_option("aos_fused", fused)
        try:
            off_g, cnt_g = ops.aos_scan(ctx, _dev(buf))
        finally:
            ctx.set_option("aos_fused", 1)
        assert cnt_g.tolist() == cnt_w.tolist(), (fused, cnt_g.tolist(), cnt_w.tolist())
        assert np.array_equal(off_g.cpu().numpy().astype(np.uint64), off_w), fused
    return off_w


def test_aos_scan_with_anomalies(ctx, oracle_mod):
    imdt, truth, imtr, aos = _downlink(n_frames=5)
    buf = synth.build_aos_file(aos, empty_every=5, bad_crc_at={2, 7, 40}, bad_inject_at={4, 33},
                               prefix=b"\x00\x11\x22" * 7, suffix=synth.AOS_SYNC + b"\x01" * 300)
    _check_aos(ctx, oracle_mod, buf)


def test_aos_scan_false_syncs(ctx, oracle_mod):
    """false sync words inside payloads (shadowed), inside a rejected frame (visited), overlapping frames"""
    imdt, truth, imtr, aos = _downlink(n_frames=4)
    aos = aos.copy()
    rng = np.random.default_rng(5)
    for i in rng.choice(aos.shape[0], 12, replace=False):
        p = int(rng.integers(20, 880))
        aos[i, p:p + 4] = np.frombuffer(synth.AOS_SYNC, np.uint8)
    crc = synth.crc16_rows(aos[:, 4:894])
    aos[:, 894], aos[:, 895] = crc >> 8, crc & 0xFF
    bad = aos.copy()
    bad[5, 700] ^= 1   # CRC failure on frames that also contain a false sync
    bad[9, 20] ^= 1
    _check_aos(ctx, oracle_mod, aos.reshape(-1))
    _check_aos(ctx, oracle_mod, bad.reshape(-1))
    # two VALID frames overlapping by less than 1024 bytes: only the first is accepted
    f = aos[3].copy()
    emb = np.concatenate([aos[0][:300], f, aos[1]])
    _check_aos(ctx, oracle_mod, emb)
    # sync words every 4 bytes (candidate table overflow path)
    storm = np.tile(np.frombuffer(synth.AOS_SYNC, np.uint8), 3000)
    _check_aos(ctx, oracle_mod, np.concatenate([storm, aos[:8].reshape(-1), storm]))


def test_aos_scan_cadence_breaks(ctx, oracle_mod):
    """the single-pass kernel assumes back-to-back frames at the phase of the first sync word: byte slips (inserted /
    dropped bytes), a file that starts with a long junk prefix, a corrupted sync word, garbage runs containing sync
    words, a false sync inside an EMPTY frame and inside a bad-CRC frame -- all bit-identical to the sequential scan"""
    imdt, truth, imtr, aos = _downlink(n_frames=12, tc=32, tl=8)
    assert aos.shape[0] > 300
    flat = aos.reshape(-1).copy()
    rng = np.random.default_rng(11)
    sync = np.frombuffer(synth.AOS_SYNC, np.uint8)
    # byte slips: 5 bytes inserted after frame 40, 3 bytes dropped inside frame 90, 1 inserted after frame 200
    a = np.concatenate([flat[:40 * 1024], rng.integers(0, 256, 5, dtype=np.uint8), flat[40 * 1024:90 * 1024 + 500],
                        flat[90 * 1024 + 503:200 * 1024], np.zeros(1, np.uint8), flat[200 * 1024:]])
    _check_aos(ctx, oracle_mod, a)
    # long junk prefix with sync words in it (some followed by a whole frame's worth of bytes, none valid), then the frames
    junk = rng.integers(0, 256, 70000, dtype=np.uint8)
    for p in (100, 5000, 5002, 33000, 69990):
        junk[p:p + 4] = sync
    _check_aos(ctx, oracle_mod, np.concatenate([junk, flat]))
    # a corrupted sync word, a false sync inside an empty frame and inside a bad-CRC frame
    b = synth.build_aos_file(aos, empty_every=9, bad_crc_at={17, 60})
    fr = b.reshape(-1, 1024) if b.size % 1024 == 0 else None
    assert fr is not None
    fr = fr.copy()
    fr[30, 1] ^= 0x40                                         # frame 30 loses its sync word
    empties = [i for i in range(fr.shape[0]) if fr[i, 5] == 0x3F and fr[i, 10] == 0xAA]
    fr[empties[2], 500:504] = sync                            # visited: the scan continues inside an empty frame
    fr[empties[3], 1021:1024] = sync[:3]                      # ... and a sync word that straddles into the next frame? (next starts 1A: not CF)
    _check_aos(ctx, oracle_mod, fr.reshape(-1))
    # tail: the file ends in the middle of a frame, and with a lone sync word
    _check_aos(ctx, oracle_mod, np.concatenate([flat[:77 * 1024 + 333], sync]))


@pytest.mark.parametrize("n", [0, 1023, 1024, 1025, 16384, 16384 + 1024, 3 * 16384 - 5, 32768 + 7, 32768 * 3 + 1024 + 9])
def test_aos_scan_sizes_and_chunk_boundaries(ctx, oracle_mod, n):
    imdt, truth, imtr, aos = _downlink(n_frames=2)
    buf = np.concatenate([np.zeros(7, np.uint8), aos.reshape(-1)])[:n]  # frames straddle the 16 KiB chunks
    _check_aos(ctx, oracle_mod, buf)


def test_imtr_deframe(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    imdt, truth, imtr, aos = _downlink(n_frames=4)
    imtr = imtr.copy()
    imtr[1, 0] ^= 0xFF
    imtr[2, 880] ^= 0xFF
    imtr[3, 9] = 0x11
    synth.refresh_imtr_crc(imtr[3:4])
    imtr[4, 300] ^= 0x01
    imtr[20, 4:8] = 0          # seq 0 -> restart rule
    synth.refresh_imtr_crc(imtr[20:21])
    buf = synth.build_aos_file(synth.aos_frames(imtr.reshape(-1)), empty_every=9)
    off = _check_aos(ctx, oracle_mod, buf)
    want, st_w = oracle_mod.imtr_deframe(buf, off)
    for runs in (1, 0):   # run-based gather (default) and the per-frame gather
        ctx.set_option("imtr_runs", runs)
        try:
            got, st_g = ops.imtr_deframe(ctx, _dev(buf), _dev(off.astype(np.int64)))
        finally:
            ctx.set_option("imtr_runs", 1)
        assert st_g.tolist() == st_w.tolist(), runs
        assert np.array_equal(got.cpu().numpy(), want), runs


def test_imtr_deframe_clean_stream(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    imdt, truth, imtr, aos = _downlink(n_frames=6, seed=8)
    buf = aos.reshape(-1)
    off, _ = oracle_mod.aos_scan(buf)
    want, st_w = oracle_mod.imtr_deframe(buf, off)
    for runs in (1, 0):
        ctx.set_option("imtr_runs", runs)
        try:
            got, st_g = ops.imtr_deframe(ctx, _dev(buf), _dev(off.astype(np.int64)))
        finally:
            ctx.set_option("imtr_runs", 1)
        assert st_g.tolist() == st_w.tolist() and st_w[1] == st_w[0], runs
        assert np.array_equal(got.cpu().numpy(), want), runs
    assert np.array_equal(want[:imdt.size], imdt)


@pytest.mark.parametrize("prefix", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("n_aos", [1, 2, 3, 33, 34, 65, 300])
def test_imtr_deframe_alignments_and_tails(ctx, oracle_mod, prefix, n_aos):
    """every byte alignment of the payload runs (file prefix of 0..5 bytes), payload counts around the 32-frame batches of a CTA,
    a cut that ends inside the last payload; a few frames damaged so that the compaction path runs as well"""
    from opticalimageprocessor_b200 import ops
    imdt, truth, imtr, aos = _downlink(n_frames=12, tc=32, tl=8, seed=21 + prefix)
    imtr = imtr[:max(1, (n_aos * 880) // 882 + 1)].copy()
    if imtr.shape[0] > 40:
        imtr[7, 333] ^= 0x10
        imtr[39, 2] ^= 0x01
    a = synth.aos_frames(imtr.reshape(-1))[:n_aos]
    buf = np.concatenate([np.full(prefix, 0x5A, np.uint8), a.reshape(-1)])
    off, _ = oracle_mod.aos_scan(buf)
    assert off.size == a.shape[0]
    want, st_w = oracle_mod.imtr_deframe(buf, off)
    for runs in (1, 0):
        ctx.set_option("imtr_runs", runs)
        try:
            got, st_g = ops.imtr_deframe(ctx, _dev(buf), _dev(off.astype(np.int64)))
        finally:
            ctx.set_option("imtr_runs", 1)
        assert st_g.tolist() == st_w.tolist(), runs
        assert np.array_equal(got.cpu().numpy(), want), runs


@pytest.mark.parametrize("tc,tl,skip,junk", [(16, 4, (), 0), (16, 4, {3}, 0), (24, 2, {2, 3}, 333), (1536, 1, (), 0)])
def test_image_frames_index_and_unpack(ctx, oracle_mod, tc, tl, skip, junk):
    from opticalimageprocessor_b200 import ops
    imdt, truth = synth.make_imdt(5, tc, tl, seed=10, skip_seqs=skip, junk_prefix=junk)
    n_w, aux_w, pan_w, mss_w, st_w = oracle_mod.image_frames(imdt, tc, tl)
    d = _dev(imdt)
    ents, st_g = ops.image_frames_index(ctx, d, tc, tl)
    assert st_g.tolist() == st_w.tolist()
    aux, pan, mss = ops.unpack_frames(ctx, d, tc, tl, ents, int(st_g[1]))
    ctx.sync()
    assert np.array_equal(aux.cpu().numpy(), aux_w)
    assert np.array_equal(pan.cpu().numpy(), pan_w)
    assert np.array_equal(mss.cpu().numpy(), mss_w)


def test_image_frames_incomplete_and_false_signature(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    tc, tl = 16, 4
    imdt, truth = synth.make_imdt(4, tc, tl, seed=9)
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    for variant in range(3):
        buf = imdt.copy()
        if variant == 0:
            buf = buf[100:]
        elif variant == 1:
            pos = frame_bytes + 192 * tl + 64
            buf[pos:pos + 4] = np.frombuffer(synth.IMG_SIG, np.uint8)
        else:
            buf = buf[:-50]  # last trailer cut
        n_w, aux_w, pan_w, mss_w, st_w = oracle_mod.image_frames(buf, tc, tl)
        d = _dev(buf)
        ents, st_g = ops.image_frames_index(ctx, d, tc, tl)
        assert st_g.tolist() == st_w.tolist(), variant
        aux, pan, mss = ops.unpack_frames(ctx, d, tc, tl, ents, int(st_g[1]))
        ctx.sync()
        assert np.array_equal(pan.cpu().numpy(), pan_w) and np.array_equal(aux.cpu().numpy(), aux_w)
        assert np.array_equal(mss.cpu().numpy(), mss_w)


@pytest.mark.parametrize("shift", [0, 1, 3, 4, 7, 13, 15, 16, 21])
def test_image_frames_index_pointer_alignment_and_partial_signatures(ctx, oracle_mod, shift):
    """the trailer search works on aligned 16-byte chunks: every alignment of the buffer pointer (a view `shift` bytes into
    an allocation), signatures straddling chunk / warp boundaries (frame lengths are 12 mod 16, so the trailers walk
    through all phases), first-two- and first-three-byte matches that are no signature"""
    from opticalimageprocessor_b200 import ops
    tc, tl = 24, 2
    imdt, truth = synth.make_imdt(9, tc, tl, seed=31 + shift)
    buf = imdt.copy()
    rng = np.random.default_rng(shift)
    sig = np.frombuffer(synth.IMG_SIG, np.uint8)
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    spots = [int(p) for p in rng.integers(0, buf.size - 8, 300) if p % frame_bytes < frame_bytes - 180]
    for p in spots[:200]:                             # decoys inside aux / pixel data (not in the trailers): EB 90, EB 90 E1
        buf[p:p + 2] = sig[:2]
    for p in spots[200:]:
        buf[p:p + 3] = sig[:3]
    n_w, aux_w, pan_w, mss_w, st_w = oracle_mod.image_frames(buf, tc, tl)
    d = _dev(np.concatenate([np.full(shift, 0xEB, np.uint8), buf]))[shift:]
    assert d.data_ptr() % 16 == shift % 16
    ents, st_g = ops.image_frames_index(ctx, d, tc, tl)
    assert st_g.tolist() == st_w.tolist()
    aux, pan, mss = ops.unpack_frames(ctx, d, tc, tl, ents, int(st_g[1]))
    ctx.sync()
    assert np.array_equal(pan.cpu().numpy(), pan_w) and np.array_equal(aux.cpu().numpy(), aux_w)
    assert np.array_equal(mss.cpu().numpy(), mss_w)


@pytest.mark.parametrize("world", [2, 5])
def test_frame_index_on_pieces_of_the_stream(ctx, oracle_mod, world):
    """oip_image_frames_hits on every piece (+ the 174-byte head of what follows; the pieces are views into one allocation,
    i.e. arbitrarily aligned pointers) and oip_image_frames_chain over the joined table == oip_image_frames_index on the
    whole stream == the oracle (the IMDT pieces stay where the sharded re-framing produced them)"""
    from opticalimageprocessor_b200 import ops, sharding
    from test_sharding_cpu import _imdt_cases, _piece_cuts
    tc, tl, frame_bytes, cases = _imdt_cases()
    for i, buf in enumerate(cases):
        st_w = oracle_mod.image_frames(buf, tc, tl)[4]
        d = _dev(buf)
        ents_1, st_1 = ops.image_frames_index(ctx, d, tc, tl)
        assert st_1.tolist() == st_w.tolist(), i
        h1, t1 = ops.image_frames_hits(ctx, d)
        assert np.all(h1[1:] > h1[:-1]) and all(bytes(t1[k, :4]) == synth.IMG_SIG for k in range(len(h1)))
        for variant in range(3):
            cuts = _piece_cuts(buf.size, frame_bytes, world, variant)
            sizes = [cuts[r + 1] - cuts[r] for r in range(world)]
            payloads = []
            for r in range(world):
                ext = d[cuts[r]:min(buf.size, cuts[r + 1] + sharding.FRAME_HALO)]      # piece + halo, a view
                payloads.append(sharding.frames_local_hits(lambda: ops.image_frames_hits(ctx, ext), sizes, r))
            ents, st = sharding.frames_chain_all(payloads, sizes, tc, tl)
            assert st.tolist() == st_w.tolist(), (i, variant)
            assert sum(len(p[0]) for p in payloads) == len(h1)
            for f in range(int(st[1])):
                assert ents[f].frame_off == ents_1[f].frame_off and list(ents[f].tile_off) == list(ents_1[f].tile_off), (i, variant, f)


def test_full_downlink_to_raw(ctx, oracle_mod):
    """AOS file -> payloads -> IMDT -> frames -> PAN/MSS/AUX, reference geometry tile width, vs ground truth"""
    from opticalimageprocessor_b200 import ops
    tc, tl = 1536, 2
    imdt, truth = synth.make_imdt(3, tc, tl, seed=12, skip_seqs={2})
    imtr = synth.imtr_frames(imdt, chid=0x11)
    aos = synth.aos_frames(imtr.reshape(-1))
    buf = synth.build_aos_file(aos, empty_every=64, bad_crc_at={10, 500})
    d = _dev(buf)
    off, cnt = ops.aos_scan(ctx, d)
    assert cnt.tolist() == [aos.shape[0], 2, len(range(0, aos.shape[0], 64))]
    got_imdt, st = ops.imtr_deframe(ctx, d, off)
    assert st[7] == 0x11
    ents, fst = ops.image_frames_index(ctx, got_imdt, tc, tl)
    assert fst.tolist() == [2, 3, 0, 3]
    aux, pan, mss = ops.unpack_frames(ctx, got_imdt, tc, tl, ents, 3)
    ctx.sync()
    pan, mss, aux = pan.cpu().numpy(), mss.cpu().numpy(), aux.cpu().numpy()
    for s in (1, 3):
        a, p, m = truth[s]
        assert np.array_equal(pan[(s - 1) * 4 * tl:s * 4 * tl], p)
        assert np.array_equal(mss[(s - 1) * tl:s * tl], m) and np.array_equal(aux[s - 1], a)
    assert not pan[4 * tl:8 * tl].any()


def test_fused_pipeline_from_frame_tiles(ctx, oracle_mod):
    """raw image frames (BE16 sub-image layout inside the IMDT stream) straight into the fused PAN kernel"""
    import ctypes as C
    from opticalimageprocessor_b200 import capi, ops
    tc, tl, n_fr, f = 64, 8, 6, 10
    W = 8 * tc
    streams, pans, tabs = [], [], []
    for i in range(2):
        imdt, truth = synth.make_imdt(n_fr, tc, tl, seed=30 + i, junk_prefix=3 * i)
        d = _dev(imdt)
        ents, st = ops.image_frames_index(ctx, d, tc, tl)
        assert st[1] == n_fr
        tab = np.array([[ents[k].tile_off[j] for j in range(40)] for k in range(n_fr)], np.int64)
        streams.append(d)
        tabs.append(_dev(tab))
        pans.append(np.concatenate([truth[s][1] for s in range(1, n_fr + 1)]))
    rows = n_fr * 4 * tl
    rng = np.random.default_rng(1)
    kbs = [synth.rrc_coeffs(W, 77 + i) for i in range(2)]
    want = oracle_mod.pan_pipeline(pans, kbs, [0, 1.37], [0, -2.61], f)
    out = torch.empty((rows, ops.pan_out_width(2, W, f)), dtype=torch.uint16, device="cuda")
    d = capi.PanDesc()
    d.n_ccd, d.w, d.total_rows, d.row0, d.n_rows = 2, W, rows, 0, rows
    d.fold_half, d.section_rows, d.row_guard = f, 30000, 32767
    keep = []
    for i in range(2):
        c = d.ccd[i]
        c.fmt, c.n_seg = capi.FMT_BE16_TILES, 1
        c.seg[0] = capi.RowSeg(streams[i].data_ptr(), 0, rows, 0)
        kb = _dev(kbs[i])
        keep.append(kb)
        c.d_kb, c.shifted, c.dX, c.dY = kb.data_ptr(), int(i > 0), [0, 1.37][i], [0, -2.61][i]
        c.d_tile_off, c.tile_cols, c.tile_lines = tabs[i].data_ptr(), tc, tl
    d.d_out, d.out_pitch_px = out.data_ptr(), out.stride(0)
    capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
    capi.check(ctx.lib.oip_pan_check_error(ctx.h))
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("tc,tl,n_fr,junk,skip,f", [(64, 8, 6, 4, (), 10), (288, 8, 5, 0, (3,), 20), (320, 64, 3, 8, (), 100),
                                                  (272, 8, 6, 3, (), 10), (1024, 16, 2, 12, (), 100), (1536, 8, 3, 4, (2,), 100)])
def test_fused_fast_path_from_frame_tiles(ctx, oracle_mod, tc, tl, n_fr, junk, skip, f):
    """K9: the fast kernel gathers its stage rows straight from the sub-images (4-byte cp.async + mbarrier), windows
    that straddle sub-image columns / rows / frames, a zero-filled gap frame, and a stream whose junk prefix leaves the
    frames unaligned (-> generic kernel); always bit-identical to the oracle on the reassembled PAN strips"""
    import ctypes as C
    from opticalimageprocessor_b200 import capi, ops
    W = 8 * tc
    streams, pans, tabs = [], [], []
    for i in range(3):
        imdt, truth = synth.make_imdt(n_fr, tc, tl, seed=40 + i, junk_prefix=junk, skip_seqs=set(skip), max_dn=65536 if i == 1 else 4096)
        d = _dev(imdt)
        ents, st = ops.image_frames_index(ctx, d, tc, tl)
        assert st[1] == n_fr
        tabs.append(ops.frame_tile_table(ents, n_fr))
        streams.append(d)
        pans.append(np.concatenate([truth[s][1] if s in truth else np.zeros((4 * tl, W), np.uint16) for s in range(1, n_fr + 1)]))
    kbs = [synth.rrc_coeffs(W, 77 + i) for i in range(3)]
    dX, dY = [0, 1.37, -0.83], [0, -2.61, 3.19]
    S, G = (30000, 32767) if n_fr * 4 * tl < 400 else (150, 160)
    want = oracle_mod.pan_pipeline(pans, kbs, dX, dY, f, S, G)
    keep = []
    got, desc = ops.pan_pipeline_from_frames(ctx, streams, tabs, tc, tl, [_dev(k) for k in kbs], dX, dY, f, section_rows=S, row_guard=G,
                                             keep=keep)
    st = (C.c_int64 * 4)()
    capi.check(ctx.lib.oip_pan_plan_coverage(C.byref(desc), 1, 128, None, st))
    if junk % 4 == 0 and tc >= 272:   # 4-byte aligned frames, a stage window spans at most two sub-image columns
        assert st[1] > 0.9 * (st[0] + st[1]), f"fast kernel share {st[1]} of {st[0] + st[1]} px"
    else:
        assert st[1] == 0
    bad = np.argwhere(got.cpu().numpy() != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


@pytest.mark.parametrize("tc,tl,n_fr,skip", [(64, 8, 6, ()), (288, 16, 4, (2,))])
def test_downlink_to_stitched_matches_oracle_chain(ctx, oracle_mod, tc, tl, n_fr, skip):
    """oip_downlink_to_stitched: AOS files in, stitched PAN raster (+ aux, MSS) out, against the oracle's chain
    aos_scan -> imtr_deframe -> image_frames -> pan_pipeline; empty frames, a corrupted duplicate and a bad inject word in
    every downlink, a zero-filled gap frame in the second case"""
    from opticalimageprocessor_b200 import ops
    W, f = 8 * tc, 12
    files, pans, auxs, msss = [], [], [], []
    for i in range(3):
        imdt, _ = synth.make_imdt(n_fr, tc, tl, seed=60 + i, skip_seqs=set(skip), max_dn=4096 if i else 65536)
        aos = synth.build_aos_file(synth.aos_frames(synth.imtr_frames(imdt, chid=0x11 + 0x11 * (i & 1)).reshape(-1)), empty_every=7 + i,
                                   bad_crc_at={3, 50 + i}, bad_inject_at={9}, prefix=b"\x00" * (5 * i))
        off, cnt = oracle_mod.aos_scan(aos)
        stream, st = oracle_mod.imtr_deframe(aos, off)
        n, aux, pan, mss, fst = oracle_mod.image_frames(stream, tc, tl)
        assert n == n_fr
        files.append(_dev(aos)); pans.append(pan); auxs.append(aux); msss.append(mss)
    kbs = [synth.rrc_coeffs(W, 90 + i) for i in range(3)]
    dX, dY = [0, 1.37, -0.83], [0, -2.61, 3.19]
    S, G = 150, 160
    want = oracle_mod.pan_pipeline(pans, kbs, dX, dY, f, S, G)
    for threads in (1, 0, 1):   # stage 1 of the three CCDs side by side (child contexts + host threads, the default; twice: the
        ctx.set_option("downlink_threads", threads)   # second call reuses the child contexts) and one after the other
        try:
            l0 = ctx.launches
            got, stats, aux_g, mss_g = ops.downlink_to_stitched(ctx, files, tc, tl, [_dev(k) for k in kbs], dX, dY, f, section_rows=S,
                                                                row_guard=G, want_aux=True, want_mss=True)
            ctx.sync()
        finally:
            ctx.set_option("downlink_threads", 1)
        assert ctx.launches > l0 + 20, threads
        assert got.shape == want.shape
        bad = np.argwhere(got.cpu().numpy() != want)
        assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()} (threads={threads})"
        for i in range(3):
            assert stats[i]["frames"][1] == n_fr and stats[i]["aos"][1] >= 1 and stats[i]["aos"][2] >= 1
            assert np.array_equal(aux_g[i].cpu().numpy(), auxs[i]) and np.array_equal(mss_g[i].cpu().numpy(), msss[i])


def test_downlink_to_stitched_reports_the_failing_ccd(ctx):
    """an error inside a worker thread (here: a JPEG-2000 compressed frame in the second downlink, which is refused) reaches
    the caller with its text and the CCD it belongs to"""
    from opticalimageprocessor_b200 import ops
    tc, tl = 16, 4
    imdt, _ = synth.make_imdt(2, tc, tl, seed=3)
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    z = imdt.copy()
    z[frame_bytes - 172 + 4] |= 0x05          # z_ratio of frame 1 (ref aux_separator.h:639)
    wrap = lambda b: _dev(synth.aos_frames(synth.imtr_frames(b, chid=0x11).reshape(-1)).reshape(-1))
    good, bad = wrap(imdt), wrap(z)
    kbs = [_dev(synth.rrc_coeffs(8 * tc, 5 + i)) for i in range(3)]
    for threads in (1, 0):
        ctx.set_option("downlink_threads", threads)
        try:
            with pytest.raises(RuntimeError, match="JPEG-2000") as ei:
                ops.downlink_to_stitched(ctx, [good, bad, good], tc, tl, kbs, [0, 1.0, -1.0], [0, 1.0, -1.0], 4)
        finally:
            ctx.set_option("downlink_threads", 1)
        if threads:
            assert "ccd 1" in str(ei.value)


@pytest.mark.parametrize("bits", [10, 12])
def test_packed_lines_extension(ctx, oracle_mod, bits):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(bits)
    w, rows, f = 512, 200, 10
    imgs = [rng.integers(0, 1 << bits, (rows, w), dtype=np.uint16) for _ in range(2)]
    raws = [synth.pack_bits(im, bits) for im in imgs]
    fmt = ops.FMT_PACK12 if bits == 12 else ops.FMT_PACK10
    for im, raw in zip(imgs, raws):
        assert np.array_equal(oracle_mod.unpack_bits(raw, bits, w, rows, raw.shape[1]), im)
        assert np.array_equal(ops.unpack_lines(ctx, _dev(raw), fmt, w).cpu().numpy(), im)
    for w2 in (496, 500, 16, 1040):  # multiples of 16 take the word path, anything else the byte path
        im = rng.integers(0, 1 << bits, (37, w2), dtype=np.uint16)
        raw = synth.pack_bits(im, bits)
        assert np.array_equal(ops.unpack_lines(ctx, _dev(raw), fmt, w2).cpu().numpy(), im)
    kbs = [synth.rrc_coeffs(w, 5 + i) for i in range(2)]
    want = oracle_mod.pan_pipeline(imgs, kbs, [0, -0.83], [0, 3.19], f)
    got = ops.pan_pipeline(ctx, [_dev(r) for r in raws], [_dev(k) for k in kbs], [0, -0.83], [0, 3.19], f, fmt=fmt, w=w)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("prefix,restart_at", [(b"", None), (b"\x00" * 13, 77)])
@pytest.mark.parametrize("world", [2, 5])
def test_stage1_byte_range_shards_on_the_gpu(ctx, oracle_mod, world, prefix, restart_at):
    """oip_aos_scan_shard + oip_imtr_deframe_shard driven by the same two exchanges as the CPU test
    (tests/test_sharding_cpu.py, where gloo carries them): shards on and off the frame cadence, a sequence restart and a bad
    IMTR frame -- payload list, counters, IMDT bytes and IMTR stats equal the sequential whole-file oracle"""
    from opticalimageprocessor_b200 import ops, sharding
    from test_sharding_cpu import _run_stage1_sharded, _stage1_file
    buf = _stage1_file(prefix=prefix, restart_at=restart_at)
    off_w, cnt_w = oracle_mod.aos_scan(buf)
    imdt_w, st_w = oracle_mod.imtr_deframe(buf, off_w)

    def aos_shard(sub, own, carry):
        o, c, co = ops.aos_scan_shard(ctx, _dev(sub), own, carry)
        return o.cpu().numpy().astype(np.uint64), c, co

    def imtr_shard(ext, offs, skip, nf):
        piece, info = ops.imtr_deframe_shard(ctx, _dev(ext), _dev(offs.astype(np.int64)), skip, nf)
        return piece.cpu().numpy(), info
    for ranges in (sharding.aos_shard_ranges(buf.size, world, len(prefix)),
                   [(buf.size * r // world + (3 if r else 0), buf.size * (r + 1) // world + (3 if r + 1 < world else 0)) for r in range(world)]):
        off, cnt, imdt, stats, carries = _run_stage1_sharded(buf, world, ranges, aos_shard, imtr_shard)
        assert np.array_equal(off, off_w) and cnt.tolist() == cnt_w.tolist()
        assert stats == st_w.tolist(), (stats, st_w.tolist())
        assert np.array_equal(imdt, imdt_w)


def test_aos_scan_without_a_single_valid_frame(ctx, oracle_mod):
    """a stretch (here: a whole file) of fill frames and corrupted frames used to be walked by one thread; now
    every candidate that no valid frame can shadow is resolved on its own.  40000 candidates, counters equal the oracle"""
    imdt, truth, imtr, aos = _downlink(n_frames=2)
    bad = np.tile(aos, (400 // aos.shape[0] + 1, 1))[:400].copy()
    bad[:, 600] ^= 0x21                                          # every frame fails its CRC
    emp = np.tile(synth.aos_empty_frame(), (100, 1))
    buf = np.concatenate([np.concatenate([bad, emp]).reshape(-1)] * 80)
    off_w, cnt_w = oracle_mod.aos_scan(buf)
    assert cnt_w.tolist() == [0, 32000, 8000]
    _check_aos(ctx, oracle_mod, buf)


@pytest.mark.parametrize("bits,w,rows", [(12, 2048, 700), (10, 2048, 700), (12, 8192, 260), (10, 4096, 300), (12, 1000, 300)])
def test_packed_samples_on_the_fast_kernel(ctx, oracle_mod, bits, w, rows):
    """MSB-first packed 12 / 10-bit lines (north_star (1); not in the reference) through the fused path: the fast kernel
    takes them -- the packed stage is unpacked in shared memory -- whenever the line is a whole number of 16-sample groups;
    3 CCDs, both signs of dY, short sections (edges, stale rows), bit-identical to the oracle on the unpacked samples"""
    import ctypes as C
    from opticalimageprocessor_b200 import capi, ops
    rng = np.random.default_rng(bits * 1000 + w)
    f = 50
    imgs = [rng.integers(0, 1 << bits, (rows, w), dtype=np.uint16) for _ in range(3)]
    raws = [synth.pack_bits(im, bits) for im in imgs]
    fmt = ops.FMT_PACK12 if bits == 12 else ops.FMT_PACK10
    kbs = [synth.rrc_coeffs(w, 5 + i) for i in range(3)]
    dX, dY = [0, 1.37, -0.83], [0, -2.61, 3.19]
    S, G = 250, 260
    want = oracle_mod.pan_pipeline(imgs, kbs, dX, dY, f, S, G)
    dev = [_dev(r) for r in raws]
    dkb = [_dev(k) for k in kbs]
    got = ops.pan_pipeline(ctx, dev, dkb, dX, dY, f, fmt=fmt, w=w, section_rows=S, row_guard=G)
    out = torch.empty_like(got)
    d = ops.make_pan_desc(dev, fmt, dkb, dX, dY, [False, True, True], f, out, section_rows=S, row_guard=G, w=w)
    st = (C.c_int64 * 4)()
    capi.check(ctx.lib.oip_pan_plan_coverage(C.byref(d), 1, 128, None, st))
    if w % 16 == 0:
        assert st[1] > 0.9 * (st[0] + st[1]), f"fast kernel share {st[1]} of {st[0] + st[1]} px"
    else:
        assert st[1] == 0
    bad = np.argwhere(got.cpu().numpy() != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


def test_mixed_source_formats_one_launch_per_class(ctx, oracle_mod):
    """a strip whose CCDs arrive in different formats (big-endian raster, packed 12-bit lines, native raster): the fast
    path runs one pan_fast_kernel launch per source-format class over its own slice of the warp-tile list; the stitched
    raster is bit-identical to the oracle on the decoded samples"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(77)
    w, rows, f = 2048, 700, 50
    imgs = [rng.integers(0, 4096, (rows, w), dtype=np.uint16) for _ in range(3)]
    dev = [_dev(imgs[0].byteswap()), _dev(synth.pack_bits(imgs[1], 12)), _dev(imgs[2])]
    fmts = [ops.FMT_BE16, ops.FMT_PACK12, ops.FMT_LE16]
    kbs = [synth.rrc_coeffs(w, 15 + i) for i in range(3)]
    dX, dY = [0, 1.37, -0.83], [0, -2.61, 3.19]
    S, G = 250, 260
    want = oracle_mod.pan_pipeline(imgs, kbs, dX, dY, f, S, G)
    l0 = ctx.launches
    got = ops.pan_pipeline(ctx, dev, [_dev(k) for k in kbs], dX, dY, f, fmt=fmts, w=w, section_rows=S, row_guard=G)
    assert ctx.launches - l0 >= 3  # generic tiles + the two classes present
    bad = np.argwhere(got.cpu().numpy() != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"
