"""Host-side planner of the fused PAN pipeline (no GPU): every output pixel is planned exactly once, by the
regular-interior fast kernel or by the exact generic kernel, for the reference geometry, BASELINE configs,
multi-section strips, multi-GPU shards with halo segments and awkward alignments."""
import ctypes as C

import numpy as np
import pytest

from opticalimageprocessor_b200 import build, capi
from opticalimageprocessor_b200.capi import PanDesc, RowSeg


@pytest.fixture(scope="module")
def lib():
    build.build()
    return capi.load()


def _desc(n, w, total, f, dX, dY, S=30000, G=32767, row0=0, n_rows=None, segs=None, base=0x7F0000000000, out_pitch=None,
          out_base=0x7E0000000000, fmt=capi.FMT_BE16):
    d = PanDesc()
    d.n_ccd, d.w, d.total_rows, d.row0 = n, w, total, row0
    d.n_rows = total - row0 if n_rows is None else n_rows
    d.fold_half, d.section_rows, d.row_guard = f, S, G
    for i in range(n):
        c = d.ccd[i]
        c.fmt = fmt
        if segs is None:
            c.n_seg = 1
            c.seg[0] = RowSeg(base + i * (1 << 36), 0, total, 2 * w)
        else:
            c.n_seg = len(segs)
            for k, (r0, nr) in enumerate(segs):
                c.seg[k] = RowSeg(base + i * (1 << 36) + k * (1 << 32), r0, nr, 2 * w)
        c.d_kb = 0x7D0000000000 + i * (1 << 24)
        c.shifted = int(i > 0)
        c.dX, c.dY = dX[i], dY[i]
    d.d_out = out_base
    out_w = n * w - 2 * (n - 1) * f
    d.out_pitch_px = out_w if out_pitch is None else out_pitch
    return d, out_w


def _coverage(lib, d, out_w, enable=1, rows=128):
    cover = np.zeros((d.n_rows, d.out_pitch_px), np.uint8)
    st = (C.c_int64 * 4)()
    capi.check(lib.oip_pan_plan_coverage(C.byref(d), enable, rows, cover.ctypes.data_as(C.c_void_p), st))
    return cover, list(st)


@pytest.mark.parametrize("n,w,total,f,S,G", [(2, 12288, 2048, 100, 30000, 32767),   # reference geometry (2 CMOS, 12288 px)
                                             (3, 8192, 1500, 100, 400, 450),         # C2 geometry, many small sections
                                             (3, 1024, 1100, 100, 400, 450), (2, 1536, 700, 100, 30000, 32767),
                                             (4, 256, 333, 3, 100, 120), (1, 520, 70, 0, 30000, 32767)])
def test_every_pixel_planned_once(lib, n, w, total, f, S, G):
    dX, dY = [0.0, 1.37, -0.83, 2.2][:n], [0.0, -2.61, 3.19, 0.4][:n]
    d, out_w = _desc(n, w, total, f, dX, dY, S, G)
    cover, st = _coverage(lib, d, out_w)
    assert set(np.unique(cover[:, :out_w])) <= {1, 2}, np.unique(cover[:, :out_w], return_counts=True)
    assert st[0] + st[1] == total * out_w
    if w >= 512 and total >= 300:
        assert st[1] > 0.6 * total * out_w, st          # most pixels are regular interior
    cover0, st0 = _coverage(lib, d, out_w, enable=0)      # fast path switched off: all generic
    assert np.all(cover0[:, :out_w] == 1) and st0[1] == 0


def test_bench_config_is_almost_entirely_fast(lib):
    """C2: 3 x 8192 px, 32768 lines, fold 200 -- only section edges / map-rounding rows stay generic"""
    d, out_w = _desc(3, 8192, 32768, 100, [0, 1.37, -0.83], [0, -2.61, 3.19])
    _, st = _coverage_stats(lib, d)
    assert st[0] + st[1] == 32768 * out_w
    assert st[1] > 0.995 * 32768 * out_w, st


def _coverage_stats(lib, d, enable=1, rows=128):
    st = (C.c_int64 * 4)()
    capi.check(lib.oip_pan_plan_coverage(C.byref(d), enable, rows, None, st))
    return None, list(st)


def test_shard_with_halo_segments(lib):
    """a middle shard of a long strip: own rows in segment 0, halo rows above / below in segments 1 and 2"""
    total, row0, n_rows = 4096, 1024, 1024
    d, out_w = _desc(3, 2048, total, 50, [0, 1.37, -0.83], [0, -2.61, 3.19], S=30000, G=32767, row0=row0, n_rows=n_rows,
                     segs=[(row0, n_rows), (row0 - 16, 16), (row0 + n_rows, 16)])
    cover, st = _coverage(lib, d, out_w)
    assert set(np.unique(cover[:, :out_w])) <= {1, 2}
    assert st[0] + st[1] == n_rows * out_w and st[1] > 0.9 * n_rows * out_w


@pytest.mark.parametrize("out_pitch,out_base,base", [(None, 0x7E0000000002, 0x7F0000000000), (3 * 1024 - 4 * 25 + 3, 0x7E0000000000, 0x7F0000000000),
                                                     (None, 0x7E0000000000, 0x7F0000000008)])
def test_unaligned_buffers_fall_back_to_generic(lib, out_pitch, out_base, base):
    d, out_w = _desc(3, 1024, 500, 25, [0, 1.37, -0.83], [0, -2.61, 3.19], out_pitch=out_pitch, out_base=out_base, base=base)
    cover, st = _coverage(lib, d, out_w)
    assert np.all(cover[:, :out_w] == 1) and st[1] == 0


def test_odd_fold_keeps_store_alignment(lib):
    """fold_half = 7: CCD 1 starts at an odd output column -> fast spans start at the next multiple of 4 / 8"""
    d, out_w = _desc(3, 1024, 400, 7, [0, 1.37, -0.83], [0, -2.61, 3.19], out_pitch=3048)   # out_w = 3044: pad the pitch to 8 px
    cover, st = _coverage(lib, d, out_w)
    assert set(np.unique(cover[:, :out_w])) <= {1, 2} and st[1] > 0.8 * 400 * out_w
    fast_cols = np.where((cover == 2).any(axis=0))[0]
    runs = np.split(fast_cols, np.where(np.diff(fast_cols) > 1)[0] + 1)
    assert all(r[0] % 4 == 0 for r in runs)


# ------------------------------------------------------------------------------------------ MSS planner
def _mss_desc(wb, lines, lps=20000, overlap=520, off=0, keep=False, min_lines=1500, scale=1.0):
    from opticalimageprocessor_b200.capi import MssDesc
    d = MssDesc()
    d.fmt, d.wb, d.lines, d.pitch_px = capi.FMT_LE16, wb, lines, 4 * wb
    for b in range(4):
        d.d_kb[b] = 0x7D0000000000 + b * (1 << 24)
        d.cX[2 * b], d.cX[2 * b + 1] = 0.8 + 0.1 * b, -1.5e-4 * (b + 1) * scale
        d.cY[3 * b], d.cY[3 * b + 1], d.cY[3 * b + 2] = -3.2 + b, 2e-4 * (b + 1) * scale, -1e-8 * (b - 1.5) * scale
    d.lines_per_section, d.line_offset, d.overlap, d.keep_leading, d.min_process_lines = lps, off, overlap, int(keep), min_lines
    return d


@pytest.mark.parametrize("wb,lines,lps,overlap,off,keep,scale", [(3072, 2400, 20000, 520, 0, False, 1.0), (96, 700, 300, 40, 0, True, 10.0),
                                                                  (512, 650, 256, 32, 10, False, 1.0), (1000, 900, 333, 17, 5, False, 3.0)])
def test_mss_every_sample_planned_once(lib, wb, lines, lps, overlap, off, keep, scale):
    d = _mss_desc(wb, lines, lps, overlap, off, keep, 64, scale)
    rows_out = lines - off - (0 if keep else overlap)
    cover = np.zeros((rows_out, wb, 4), np.uint8)
    st = (C.c_int64 * 4)()
    capi.check(lib.oip_mss_plan_coverage(C.byref(d), 1, 128, cover.ctypes.data_as(C.c_void_p), st))
    planned = cover[: (st[0] + st[1]) // (4 * wb)]
    assert (st[0] + st[1]) % (4 * wb) == 0 and set(np.unique(planned)) <= {1, 2}, np.unique(planned, return_counts=True)
    assert not cover[planned.shape[0]:].any()            # rows the reference never writes stay untouched
    if wb >= 512:
        assert st[1] > 0.5 * (st[0] + st[1]), list(st)
    st0 = (C.c_int64 * 4)()
    capi.check(lib.oip_mss_plan_coverage(C.byref(d), 0, 128, None, st0))
    assert st0[1] == 0 and st0[0] == st[0] + st[1]


def test_mss_reference_geometry_is_mostly_fast(lib):
    """C3-like: 3072-px bands, 20000-line sections: only the power-of-two row zones and band edges stay generic"""
    d = _mss_desc(3072, 40000)
    st = (C.c_int64 * 4)()
    capi.check(lib.oip_mss_plan_coverage(C.byref(d), 1, 128, None, st))
    assert st[1] > 0.98 * (st[0] + st[1]), list(st)


def test_mixed_source_formats_are_planned_once_per_pixel(lib):
    """CCDs of different source-format classes in one strip (big-endian raster, packed 12-bit lines, native raster): the
    warp-tile list is laid out class by class (one pan_fast_kernel launch each); every output pixel is still covered
    exactly once and the packed CCD stays on the fast path"""
    n, w, total, f = 3, 2048, 700, 50
    d, out_w = _desc(n, w, total, f, [0.0, 1.37, -0.83], [0.0, -2.61, 3.19], 250, 260)
    d.ccd[1].fmt = capi.FMT_PACK12
    d.ccd[1].seg[0] = RowSeg(0x7F0000000000 + (1 << 36), 0, total, w * 12 // 8)
    d.ccd[2].fmt = capi.FMT_LE16
    cover, st = _coverage(lib, d, out_w)
    assert set(np.unique(cover[:, :out_w])) <= {1, 2}
    assert st[0] + st[1] == total * out_w
    assert st[1] > 0.8 * total * out_w, st


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_geometries_are_planned_exactly_once(lib, seed):
    """300 random strips per seed -- 1..4 CCDs, widths off every alignment, folds, section / guard sizes, shifts with 0..3
    decimals (integer shifts included), both 16-bit byte orders, warp-tile heights, whole strips and row shards: every
    output pixel of the shard is planned exactly once (fast or generic), whatever the split between the two kernels"""
    rng = np.random.default_rng(seed)
    for it in range(300):
        n = int(rng.integers(1, 5))
        w = int(rng.choice([264, 520, 1000, 1024, 1536, 2048, 3000]))
        f = int(rng.integers(0, min(120, w // 4)))
        G = int(rng.integers(60, 600))
        S = int(rng.integers(8, G + 1))
        total = int(rng.integers(40, 1500))
        dX = [0.0] + [float(np.round(rng.uniform(-6, 6), int(rng.integers(0, 4)))) for _ in range(n - 1)]
        dY = [0.0] + [float(np.round(rng.uniform(-6, 6), int(rng.integers(0, 4)))) for _ in range(n - 1)]
        fmt = int(rng.choice([capi.FMT_BE16, capi.FMT_LE16]))
        rows = int(rng.choice([16, 61, 128, 256]))
        row0, n_rows = 0, None
        if rng.random() < 0.3 and total > 200:
            row0 = int(rng.integers(0, total // 2))
            n_rows = int(rng.integers(20, total - row0))
        d, out_w = _desc(n, w, total, f, dX, dY, S, G, row0=row0, n_rows=n_rows, fmt=fmt)
        cover, st = _coverage(lib, d, out_w, rows=rows)
        what = (seed, it, n, w, total, f, S, G, dX, dY, row0, n_rows, rows)
        assert set(np.unique(cover[:, :out_w])) <= {1, 2}, what
        assert st[0] + st[1] == d.n_rows * out_w, what


def test_mss_random_geometries_are_planned_exactly_once(lib):
    """150 random band-alignment jobs -- band widths, section lengths, overlaps, line offsets, keep-leading on and off,
    polynomials from flat to steep (x 40), warp-tile heights: every sample the reference writes is planned exactly once and
    the rows it never writes stay untouched"""
    rng = np.random.default_rng(7)
    for it in range(150):
        wb = int(rng.choice([96, 200, 512, 1000, 1536, 3072]))
        lps = int(rng.integers(100, 900))
        overlap = int(rng.integers(0, min(lps // 2, 120)))
        off = int(rng.integers(0, 30))
        keep = bool(rng.random() < 0.4)
        lines = int(rng.integers(lps + overlap + off + 10, 2500))
        scale = float(rng.choice([0.0, 0.3, 1.0, 3.0, 10.0, 40.0]))
        rows = int(rng.choice([16, 61, 128]))
        d = _mss_desc(wb, lines, lps, overlap, off, keep, int(rng.integers(1, 200)), scale)
        rows_out = lines - off - (0 if keep else overlap)
        cover = np.zeros((rows_out, wb, 4), np.uint8)
        st = (C.c_int64 * 4)()
        capi.check(lib.oip_mss_plan_coverage(C.byref(d), 1, rows, cover.ctypes.data_as(C.c_void_p), st))
        what = (it, wb, lines, lps, overlap, off, keep, scale, rows, d.min_process_lines)
        planned = cover[: (st[0] + st[1]) // (4 * wb)]
        assert (st[0] + st[1]) % (4 * wb) == 0 and set(np.unique(planned)) <= {1, 2}, what
        assert not cover[planned.shape[0]:].any(), what
