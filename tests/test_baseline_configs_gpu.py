"""BASELINE.json `configs`, one test each, at sizes the CPU oracle finishes in seconds (the bench line is C2 at full size).
Everything goes through the C ABI; integer / byte / index work bit-exact, and so is the float resampling."""
import ctypes as C

import numpy as np
import pytest
import torch

from opticalimageprocessor_b200 import synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_c1_single_ccd_4096_framed_auxsep_crc_rrc(ctx, oracle_mod):
    """C1: single-CCD 4096 px x 4096 lines, framed downlink: aux separation + CRC + radiometric correction.
    (the 12-bit packed variant of the same lines is the extension format, checked at the end)"""
    from opticalimageprocessor_b200 import ops
    tc, tl = 512, 256                                  # 8 x 512 = 4096 px, 4 x 256 = 1024 PAN lines per frame
    imdt, truth = synth.make_imdt(4, tc, tl, seed=41)  # 4 frames = 4096 PAN lines
    aos = synth.aos_frames(synth.imtr_frames(imdt, chid=0x22).reshape(-1))
    buf = synth.build_aos_file(aos, empty_every=64, bad_crc_at=set(range(7, aos.shape[0], 1024)))
    d = _dev(buf)
    off_w, cnt_w = oracle_mod.aos_scan(buf)
    off, cnt = ops.aos_scan(ctx, d)
    assert cnt.tolist() == cnt_w.tolist() and np.array_equal(off.cpu().numpy().astype(np.uint64), off_w)
    want_imdt, st_w = oracle_mod.imtr_deframe(buf, off_w)
    got_imdt, st = ops.imtr_deframe(ctx, d, off)
    assert st.tolist() == st_w.tolist() and np.array_equal(got_imdt.cpu().numpy(), want_imdt)
    n_w, aux_w, pan_w, mss_w, fst_w = oracle_mod.image_frames(want_imdt, tc, tl)
    ents, fst = ops.image_frames_index(ctx, got_imdt, tc, tl)
    assert fst.tolist() == fst_w.tolist() and fst[1] == 4
    aux, pan, mss = ops.unpack_frames(ctx, got_imdt, tc, tl, ents, 4)
    assert np.array_equal(aux.cpu().numpy(), aux_w) and np.array_equal(mss.cpu().numpy(), mss_w)
    assert pan.shape == (4096, 4096) and np.array_equal(pan.cpu().numpy(), pan_w)
    kb = synth.rrc_coeffs(4096, 43)
    want = oracle_mod.rrc(pan_w, kb)
    got = ops.inplace_rrc(ctx, pan, _dev(kb)).cpu().numpy()
    assert np.array_equal(got, want)
    # 12-bit packed lines of the same image -> unpack -> RRC
    packed = synth.pack_bits(pan_w, 12)
    unp = ops.unpack_lines(ctx, _dev(packed), ops.FMT_PACK12, 4096)
    assert np.array_equal(unp.cpu().numpy(), pan_w)
    assert np.array_equal(ops.inplace_rrc(ctx, unp, _dev(kb)).cpu().numpy(), want)


def test_c2_three_ccd_8192_strip(ctx, oracle_mod):
    """C2 geometry: 3 x 8192 px, fold 200, BE16 raw, RRC + cubic shift + stitch (3000 lines here, 32768 in bench.py)"""
    from opticalimageprocessor_b200 import ops
    rows, w, f = 3000, 8192, 100
    ccds = [synth.strip_dn(w, rows, 50 + i) for i in range(3)]
    kbs = [synth.rrc_coeffs(w, 60 + i) for i in range(3)]
    dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
    want = oracle_mod.pan_pipeline(ccds, kbs, dX, dY, f)
    got = ops.pan_pipeline(ctx, [_dev(c.byteswap()) for c in ccds], [_dev(k) for k in kbs], dX, dY, f, fmt=ops.FMT_BE16).cpu().numpy()
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


def test_c3_multispectral_three_ccd_band_alignment_and_stitch(ctx, oracle_mod):
    """C3: 4-band MSS strips of 3 CCDs (4 x 2048 px each), band alignment per CCD, then the 4-channel stitch (fold 50)"""
    from opticalimageprocessor_b200 import ops
    lines, wb, f = 2200, 2048, 25
    rng = np.random.default_rng(61)
    cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
    cY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8 * (b - 1.5)] for b in range(4)]
    kw = dict(lines_per_section=20000, line_offset=0, overlap=520, keep_leading=False, min_process_lines=1500)
    aligned_w, aligned_g = [], []
    for c in range(3):
        mixed = rng.integers(0, 4096, (lines, 4 * wb), dtype=np.uint16)
        kbs = [synth.rrc_coeffs(wb, 70 + 4 * c + b) for b in range(4)]
        planes = [oracle_mod.rrc(p, k) for p, k in zip(oracle_mod.mss_split(mixed), kbs)]
        n_w, a_w = oracle_mod.band_align(planes, cX, cY, **kw)
        n_g, a_g = ops.band_align(ctx, _dev(mixed), wb, [_dev(k) for k in kbs], cX, cY, **kw)
        assert n_g == n_w == lines - 520
        assert np.array_equal(a_g.cpu().numpy()[:n_g], a_w[:n_w])
        aligned_w.append(a_w[:n_w])
        aligned_g.append(a_g[:n_g])
    want = oracle_mod.stitch_concat_c4(aligned_w, f, None)
    got = ops.stitch_tiff_geometry(ctx, aligned_g, f).cpu().numpy()
    assert got.shape == (lines - 520, 3 * wb - 4 * f, 4) and np.array_equal(got, want)


def test_c4_scanline_block_shards_equal_whole_strip(ctx, oracle_mod):
    """C4: a 24576-px (3 x 8192) strip cut into scanline blocks -- every shard, computed on its own through
    row0 / n_rows (global section geometry), equals the same rows of the whole-strip result; the whole-strip
    result equals the oracle.  Multi-GPU with halo rows over NVLink: tests/test_multi_gpu.py."""
    from opticalimageprocessor_b200 import ops
    rows, w, f, S, G = 2600, 8192, 100, 1000, 1100      # 3 sections + stale rows inside the strip
    ccds = [synth.strip_dn(w, rows, 80 + i) for i in range(3)]
    kbs = [synth.rrc_coeffs(w, 90 + i) for i in range(3)]
    dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
    want = oracle_mod.pan_pipeline(ccds, kbs, dX, dY, f, S, G)
    dev = [_dev(c) for c in ccds]
    dkb = [_dev(k) for k in kbs]
    whole = ops.pan_pipeline(ctx, dev, dkb, dX, dY, f, section_rows=S, row_guard=G)
    assert np.array_equal(whole.cpu().numpy(), want)
    out_w = ops.pan_out_width(3, w, f)
    bounds = [0, 517, 1301, 2048, rows]
    for a, b in zip(bounds, bounds[1:]):
        out = torch.zeros((b - a, out_w), dtype=torch.uint16, device="cuda")
        d = ops.make_pan_desc(dev, ops.FMT_LE16, dkb, dX, dY, [0, 1, 1], f, out, total_rows=rows, row0=a, n_rows=b - a,
                              section_rows=S, row_guard=G)
        from opticalimageprocessor_b200 import capi
        capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
        capi.check(ctx.lib.oip_pan_check_error(ctx.h))
        assert torch.equal(out.view(torch.int16), whole[a:b].view(torch.int16)), (a, b)


def test_c5_sizes_sweep_is_size_independent(ctx):
    """C5 (throughput sweep): the result of a strip does not depend on how many lines are processed at once --
    checksum of checksums over growing inputs (property test, no oracle at these sizes)"""
    from opticalimageprocessor_b200 import ops
    w, f = 2048, 20
    rows_all = 40000                                     # > 32767: two 30000-row sections, partial last section
    g = torch.Generator(device="cuda").manual_seed(5)
    ccds = [torch.randint(0, 4096, (rows_all, w), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16) for _ in range(2)]
    kbs = [_dev(synth.rrc_coeffs(w, 95 + i)) for i in range(2)]
    dX, dY = [0.0, 1.37], [0.0, -2.61]
    whole = ops.pan_pipeline(ctx, ccds, kbs, dX, dY, f)
    ref = whole.view(torch.int16).to(torch.int64).sum(dim=1)
    out_w = ops.pan_out_width(2, w, f)
    for step in (4096, 16384):
        sums = []
        for a in range(0, rows_all, step):
            b = min(rows_all, a + step)
            out = torch.zeros((b - a, out_w), dtype=torch.uint16, device="cuda")
            d = ops.make_pan_desc(ccds, ops.FMT_LE16, kbs, dX, dY, [0, 1], f, out, total_rows=rows_all, row0=a, n_rows=b - a)
            from opticalimageprocessor_b200 import capi
            capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
            sums.append(out.view(torch.int16).to(torch.int64).sum(dim=1))
        capi.check(ctx.lib.oip_pan_check_error(ctx.h))
        assert torch.equal(torch.cat(sums), ref), step
