"""Oracle of SURVEY 8(f) N1 (Stitcher::CalcSttParameters, ref stitcher.h:148-201): the numpy restatement of
cv::phaseCorrelate against the real cv2.phaseCorrelate (the reference's own library call, stitcher.h:180).
Floating point: |d| <= 2e-3 px and |response| <= 1e-3 (cv2 runs its DFT in float32)."""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


def _pair(rows, cols, dx, dy, seed):
    """two views of one smooth random scene, the second displaced by (dx, dy) (cubic resampling), as u16 DN"""
    if rows > 16384:  # cv2.remap refuses images taller than 32767 rows: independent scenes stacked
        parts = [_pair(min(16384, rows - r), cols, dx, dy, seed + 1 + r) for r in range(0, rows, 16384)]
        return np.vstack([p[0] for p in parts]), np.vstack([p[1] for p in parts])
    rng = np.random.default_rng(seed)
    big = cv2.GaussianBlur(rng.random((rows + 64, cols + 64)).astype(np.float32), (0, 0), 1.2)
    big = (big - big.min()) / (big.max() - big.min()) * 3000 + 200
    xs, ys = np.meshgrid(np.arange(cols, dtype=np.float32) + 32, np.arange(rows, dtype=np.float32) + 32)
    a = cv2.remap(big, xs, ys, cv2.INTER_CUBIC)
    b = cv2.remap(big, xs - np.float32(dx), ys - np.float32(dy), cv2.INTER_CUBIC)
    return np.round(a).astype(np.uint16), np.round(b).astype(np.uint16)


@pytest.mark.parametrize("n", [1, 7, 100, 121, 200, 16000, 16001, 30000])
def test_optimal_dft_size(n):
    assert oracle.optimal_dft_size(n) == cv2.getOptimalDFTSize(n)


@pytest.mark.parametrize("rows,cols,dx,dy", [(400, 200, 1.37, -2.61), (1000, 180, -0.83, 3.19), (500, 96, 0.0, 0.0), (750, 200, 4.5, 7.25)])
def test_phase_correlate_matches_cv2(rows, cols, dx, dy):
    a, b = _pair(rows, cols, dx, dy, seed=rows + cols)
    (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    ox, oy, orr = oracle.phase_correlate(a.astype(np.float32), b.astype(np.float32))
    assert abs(ox - cx) <= 2e-3 and abs(oy - cy) <= 2e-3 and abs(orr - cr) <= 1e-3
    assert abs(ox - dx) < 0.5 and abs(oy - dy) < 0.5  # and it does measure the displacement (5x5 centroid: biased towards the integer peak)


@pytest.mark.parametrize("rows,cols", [(125, 125), (243, 100), (250, 125), (600, 75), (135, 81)])
def test_phase_correlate_odd_dft_sizes_match_cv2(rows, cols):
    """odd optimal DFT sizes: fftShift is the circular shift by (M // 2, N // 2), the centre is (N / 2.0, M / 2.0)"""
    assert oracle.optimal_dft_size(rows) % 2 or oracle.optimal_dft_size(cols) % 2
    a, b = _pair(rows, cols, 1.37, -2.61, seed=rows * 7 + cols)
    (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
    ox, oy, orr = oracle.phase_correlate(a.astype(np.float32), b.astype(np.float32))
    assert abs(ox - cx) <= 2e-3 and abs(oy - cy) <= 2e-3 and abs(orr - cr) <= 1e-3


def test_stt_parameters_follow_the_reference_loop():
    lines, w, ov = 4096, 512, 200
    scene_a, scene_b = _pair(lines, ov, 1.37, -2.61, seed=3)
    pan1 = np.zeros((lines, w), np.uint16); pan2 = np.zeros((lines, w), np.uint16)
    pan1[:, w - ov:] = scene_a
    pan2[:, :ov] = scene_b
    rows, mean = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, sections=4, lines_per_section=600)
    gap = (lines - 4 * 600) // 5
    assert [r[0] for r in rows] == [gap + i * (gap + 600) for i in range(4)]
    # the same loop on top of the real cv2.phaseCorrelate
    def cvcorr(s1, s2):
        (x, y), r = cv2.phaseCorrelate(s1, s2)
        return x, y, r
    rows_cv, mean_cv = oracle.stt_parameters(pan1, pan2, overlap_cols=ov, sections=4, lines_per_section=600, correlate=cvcorr)
    assert [r[4] for r in rows] == [r[4] for r in rows_cv]
    assert mean is not None and all(abs(p - q) <= 2e-3 for p, q in zip(mean, mean_cv))
    assert abs(mean[0] - 1.37) < 0.5 and abs(mean[1] + 2.61) < 0.5


@pytest.mark.parametrize("shape,dst", [((100, 77), (400, 308)), ((50, 64), (200, 256)), ((40, 31), (160, 125))])
def test_resize_cubic_matches_cv2(shape, dst):
    src = np.random.default_rng(4).random(shape).astype(np.float32) * 4000
    want = cv2.resize(src, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
    got = oracle.resize_cubic(src, dst[0], dst[1])
    assert np.abs(got - want).max() <= 1e-2  # float32 rounding (different summation order) on values up to 4000: 2.5e-6 relative


def _mss_scene(lines_pan, w, shifts, seed):
    """a PAN strip and 4 quarter-resolution bands of the same scene, band b displaced by shifts[b] PAN pixels"""
    rng = np.random.default_rng(seed)
    big = cv2.GaussianBlur(rng.random((lines_pan + 64, w + 64)).astype(np.float32), (0, 0), 3.0)
    big = (big - big.min()) / (big.max() - big.min()) * 3000 + 200
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32) + 32, np.arange(lines_pan, dtype=np.float32) + 32)
    pan = np.round(cv2.remap(big, xs, ys, cv2.INTER_CUBIC)).astype(np.uint16)
    bands = []
    for dx, dy in shifts:
        full = cv2.remap(big, xs - np.float32(dx), ys - np.float32(dy), cv2.INTER_CUBIC)
        bands.append(np.round(cv2.resize(full, (w // 4, lines_pan // 4), interpolation=cv2.INTER_AREA)).astype(np.uint16))
    return pan, bands


def test_inter_band_correlation_on_cv2_building_blocks():
    """the restated resize + phase correlation give the same shifts and polynomials as the loop on cv2's own functions"""
    pan, bands = _mss_scene(1600, 2048, [(2.0, -3.0), (1.0, 4.0), (-2.5, 1.5), (0.5, -0.5)], seed=8)
    kw = dict(slices=8, sections=2, corr_lines=640)

    def cvcorr(a, b):
        (x, y), r = cv2.phaseCorrelate(a, b)
        return x, y, r
    ours = oracle.inter_band_correlation(pan, bands, **kw)
    ref = oracle.inter_band_correlation(pan, bands, correlate=cvcorr,
                                        resize=lambda s, r, c: cv2.resize(s, (c, r), interpolation=cv2.INTER_CUBIC), **kw)
    for b in range(4):
        for s, t in zip(ours[0][b], ref[0][b]):
            assert s[3] == t[3] and abs(s[0] - t[0]) <= 2e-3 and abs(s[1] - t[1]) <= 2e-3 and abs(s[2] - t[2]) <= 1e-3
        assert np.allclose(ours[1][b], ref[1][b], atol=2e-3) and np.allclose(ours[2][b], ref[2][b], atol=2e-3)
