"""SURVEY 8(f) N1 on the GPU: oip_phase_correlate_u16 / oip_stt_parameters against the reference's own library call
(cv2.phaseCorrelate, ref stitcher.h:180) and the CalcSttParameters loop restated on top of it (oracle.stt_parameters).
Floating point (different DFT implementation): |d| <= 2e-3 px, |response| <= 1e-3."""
import numpy as np
import pytest
import torch

import oracle
from opticalimageprocessor_b200 import capi, ops
from test_phasecorr_cpu import _mss_scene, _pair

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return ops.Context(0)


@pytest.mark.parametrize("rows,cols,dx,dy", [(400, 200, 1.37, -2.61), (1000, 180, -0.83, 3.19), (500, 96, 0.0, 0.0),
                                             (750, 200, 4.5, 7.25), (16000, 200, 1.37, -2.61)])
def test_phase_correlate_matches_cv2(ctx, rows, cols, dx, dy):
    a, b = _pair(rows, cols, dx, dy, seed=rows + cols)
    (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    gx, gy, gr = ops.phase_correlate(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert abs(gx - cx) <= 2e-3 and abs(gy - cy) <= 2e-3 and abs(gr - cr) <= 1e-3, ((gx, gy, gr), (cx, cy, cr))


def test_phase_correlate_on_column_views(ctx):
    a, b = _pair(600, 200, -2.25, 1.5, seed=9)
    wide1 = np.zeros((600, 512), np.uint16); wide1[:, 312:] = a
    wide2 = np.zeros((600, 512), np.uint16); wide2[:, :200] = b
    (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    t1, t2 = torch.from_numpy(wide1).cuda(), torch.from_numpy(wide2).cuda()
    gx, gy, gr = ops.phase_correlate(ctx, t1[:, 312:], t2[:, :200])
    assert abs(gx - cx) <= 2e-3 and abs(gy - cy) <= 2e-3 and abs(gr - cr) <= 1e-3


@pytest.mark.parametrize("rows,cols", [(125, 125), (243, 100), (250, 125), (600, 75), (135, 81)])
def test_odd_dft_sizes_match_cv2(ctx, rows, cols):
    """getOptimalDFTSize gives an odd size (125, 243, 75, 135, 81 ...): OpenCV's fftShift is then the circular shift by
    (M // 2, N // 2) and the centre is (N / 2.0, M / 2.0) (e.g. `prestitch --stitch-overlap 125`)"""
    assert oracle.optimal_dft_size(rows) % 2 or oracle.optimal_dft_size(cols) % 2
    a, b = _pair(rows, cols, 1.37, -2.61, seed=rows * 7 + cols)
    (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    gx, gy, gr = ops.phase_correlate(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert abs(gx - cx) <= 2e-3 and abs(gy - cy) <= 2e-3 and abs(gr - cr) <= 1e-3, ((gx, gy, gr), (cx, cy, cr))


def test_edge_cols_equal_to_half_the_overlap(ctx):
    """`prestitch -e 100` with the default overlap 200 passes the reference CLI check (ref main.cpp:135-141) and
    correlates colRange(W-200, W-100) with colRange(100, 200) (ref stitcher.h:175-176)"""
    lines, w, ov = 4096, 512, 200
    pan1, pan2 = _strips(lines, w, ov, 1.37, -2.61, 11)

    def cvcorr(s1, s2):
        (x, y), r = cv2.phaseCorrelate(s1, s2)
        return x, y, r
    kw = dict(overlap_cols=ov, edge_cols=100, sections=3, lines_per_section=1000, threshold=0.0)
    rows_cv, mean_cv = oracle.stt_parameters(pan1, pan2, correlate=cvcorr, **kw)
    rows, mean = ops.calc_stt_parameters(ctx, torch.from_numpy(pan1).cuda(), torch.from_numpy(pan2).cuda(), **kw)
    for g, c in zip(rows, rows_cv):
        assert g[0] == c[0] and abs(g[1] - c[1]) <= 2e-3 and abs(g[2] - c[2]) <= 2e-3 and abs(g[3] - c[3]) <= 1e-3
    with pytest.raises(capi.OipError):
        ops.calc_stt_parameters(ctx, torch.from_numpy(pan1).cuda(), torch.from_numpy(pan2).cuda(), overlap_cols=ov, edge_cols=ov)


def _strips(lines, w, ov, dx, dy, seed):
    sa, sb = _pair(lines, ov, dx, dy, seed=seed)
    rng = np.random.default_rng(seed)
    pan1 = rng.integers(100, 4000, (lines, w)).astype(np.uint16)
    pan2 = rng.integers(100, 4000, (lines, w)).astype(np.uint16)
    pan1[:, w - ov:] = sa
    pan2[:, :ov] = sb
    return pan1, pan2


def test_calc_stt_parameters_matches_reference_loop(ctx):
    lines, w, ov = 8192, 1024, 200
    pan1, pan2 = _strips(lines, w, ov, 1.37, -2.61, 5)
    # make one section fail the response threshold: unrelated noise in its overlap
    rows_cv, mean_cv = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, edge_cols=4, sections=5, lines_per_section=1200)
    off_bad = rows_cv[2][0]
    pan2[off_bad:off_bad + 1200, :ov] = np.random.default_rng(1).integers(100, 4000, (1200, ov)).astype(np.uint16)

    def cvcorr(s1, s2):
        (x, y), r = cv2.phaseCorrelate(s1, s2)
        return x, y, r
    rows_cv, mean_cv = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, edge_cols=4, sections=5, lines_per_section=1200, correlate=cvcorr)
    rows, mean = ops.calc_stt_parameters(ctx, torch.from_numpy(pan1).cuda(), torch.from_numpy(pan2).cuda(), overlap_cols=ov,
                                         edge_cols=4, sections=5, lines_per_section=1200)
    assert [r[0] for r in rows] == [r[0] for r in rows_cv]
    assert [r[4] for r in rows] == [int(r[4]) for r in rows_cv] and rows[2][4] == 0
    for g, c in zip(rows, rows_cv):
        if c[4]:
            assert abs(g[1] - c[1]) <= 2e-3 and abs(g[2] - c[2]) <= 2e-3 and abs(g[3] - c[3]) <= 1e-3
    assert all(abs(p - q) <= 2e-3 for p, q in zip(mean, mean_cv))


def test_shards_add_up_to_the_whole(ctx):
    """two scanline-block shards: every section lies in exactly one of them here, and the summed sums give the same mean"""
    lines, w, ov = 8000, 512, 200
    pan1, pan2 = _strips(lines, w, ov, -0.83, 3.19, 6)
    t1, t2 = torch.from_numpy(pan1).cuda(), torch.from_numpy(pan2).cuda()
    kw = dict(overlap_cols=ov, sections=4, lines_per_section=1000)
    rows, mean = ops.calc_stt_parameters(ctx, t1, t2, **kw)
    tot = np.zeros(4)
    seen = []
    for r0, r1 in [(0, 4000), (4000, 8000)]:
        rs, _ = ops.calc_stt_parameters(ctx, t1[r0:r1], t2[r0:r1], total_lines=lines, row0=r0, **kw)
        for s in rs:
            if s[4] >= 0:
                seen.append(s[0])
                if s[4] == 1:
                    tot += [s[1], s[2], s[3], 1]
    assert sorted(seen) == [r[0] for r in rows]
    assert all(abs(tot[k] / tot[3] - mean[k]) < 1e-9 for k in range(3))


# ------------------------------------------------------------------------------------------------ N2
def test_inter_band_correlation_matches_reference_loop(ctx):
    """oip_inter_band_correlation against CalcInterBandCorrelation restated on cv2.resize + cv2.phaseCorrelate +
    numpy.polyfit (ref preproc.h:224-347, :492-550): shifts within 3e-3 px, polynomial values within 3e-3 px over the line"""
    true = [(2.0, -3.0), (1.0, 4.0), (-2.5, 1.5), (0.5, -0.5)]
    pan, bands = _mss_scene(1600, 2048, true, seed=8)
    mss = np.ascontiguousarray(np.hstack(bands))
    kw = dict(slices=8, sections=2, threshold=0.3)

    def cvcorr(a, b):
        (x, y), r = cv2.phaseCorrelate(a, b)
        return x, y, r
    ref = oracle.inter_band_correlation(pan, bands, corr_lines=640, correlate=cvcorr,
                                        resize=lambda s, r, c: cv2.resize(s, (c, r), interpolation=cv2.INTER_CUBIC), **kw)
    got = ops.calc_inter_band_correlation(ctx, torch.from_numpy(pan).cuda(), torch.from_numpy(mss).cuda(), correlation_lines=640, **kw)
    xs = np.array([0.0, 1024.0, 2047.0]) * 4
    for b in range(4):
        for g, r in zip(got[0][b], ref[0][b]):
            assert g[3] == r[3] and abs(g[2] - r[2]) <= 1e-3
            if r[2] >= 0.3:
                assert abs(g[0] - r[0]) <= 3e-3 and abs(g[1] - r[1]) <= 3e-3, (b, g, r)
            else:
                assert np.isnan(g[0]) and np.isnan(g[1])
        px = lambda c, x: sum(ck * x ** k for k, ck in enumerate(c))
        assert all(abs(px(got[1][b], x) - px(ref[1][b], x)) <= 3e-3 for x in xs / 4)
        assert all(abs(px(got[2][b], x) - px(ref[2][b], x)) <= 3e-3 for x in xs / 4)


def test_inter_band_correlation_argument_errors(ctx):
    pan = torch.zeros((1600, 2048), dtype=torch.uint16, device="cuda")
    mss = torch.zeros((400, 2048), dtype=torch.uint16, device="cuda")
    with pytest.raises(capi.OipError, match="at lease 8 slice needed"):
        ops.calc_inter_band_correlation(ctx, pan, mss, slices=4)
    with pytest.raises(capi.OipError, match="too many sections"):
        ops.calc_inter_band_correlation(ctx, pan, mss, slices=8, sections=5)
    with pytest.raises(capi.OipError, match="Not enough valid correlation values for band#1"):
        ops.calc_inter_band_correlation(ctx, pan, mss, slices=8, sections=1)


def test_estimate_tolerance_against_the_quantised_map(ctx):
    """the 2e-3 px tolerance feeds a map that cv::remap quantises to 1/32 px -- cvRound(float(x + dX) * 32), ref
    stitcher.h:96-97.  Measured (tools/stt_tolerance.py, profiles/r02_stt_tolerance.json, 2000 random shifts): the estimate is
    within 6.4e-6 px of cv2.phaseCorrelate (median 2.3e-7) and the quantised map of a 12288-px line / 30000-row section did not
    change in a single run.  Here: 150 shifts, the difference stays below 5e-5 px and at most 2 % of the runs change any
    map entry."""
    rng = np.random.default_rng(77)
    changed, worst = 0, 0.0

    def fixed(n, d):
        return np.rint((np.arange(n, dtype=np.float64) + d).astype(np.float32) * np.float32(32)).astype(np.int64)
    n = 150
    for i in range(n):
        dx, dy = rng.uniform(-4, 4), rng.uniform(-6, 6)
        a, b = _pair(1024, 200, dx, dy, seed=500 + i)
        (cx, cy), _ = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
        gx, gy, _ = ops.phase_correlate(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        worst = max(worst, abs(gx - cx), abs(gy - cy))
        changed += int((fixed(12288, gx) != fixed(12288, cx)).any() or (fixed(30000, gy) != fixed(30000, cy)).any())
    assert worst < 5e-5, worst
    assert changed <= 0.02 * n, changed
