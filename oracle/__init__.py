"""ctypes binding of the CPU oracle (oracle/liboip_oracle.so) and of the reference-backed
checkers under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF_CRC = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class FrameGeom(C.Structure):
    _fields_ = [("tile_cols", C.c_int), ("tile_lines", C.c_int)]


def build(force: bool = False) -> None:
    """compile the oracle (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboip_oracle.so")
    src = os.path.join(_HERE, "oip_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboip_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/OpticalImageProcessor"):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboip_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oipo_crc16.restype = C.c_uint16
        L.oipo_crc16.argtypes = [u8p, C.c_size_t]
        L.oipo_aos_validate.restype = C.c_int
        L.oipo_aos_validate.argtypes = [u8p] + [C.POINTER(C.c_uint32)] * 4
        L.oipo_aos_scan.restype = C.c_int64
        L.oipo_aos_scan.argtypes = [u8p, C.c_size_t, u64p, C.c_size_t, i64p]
        L.oipo_aos_scan_range.restype = C.c_int64
        L.oipo_aos_scan_range.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_size_t, u64p, C.c_size_t, i64p, C.POINTER(C.c_uint64)]
        L.oipo_imtr_deframe.restype = C.c_int64
        L.oipo_imtr_deframe.argtypes = [u8p, u64p, C.c_int64, u8p, C.c_size_t, i64p]
        L.oipo_image_frames.restype = C.c_int64
        L.oipo_image_frames.argtypes = [u8p, C.c_size_t, C.POINTER(FrameGeom), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int64, i64p]
        L.oipo_rrc_u16.restype = None
        L.oipo_rrc_u16.argtypes = [u16p, C.c_int, C.c_int64, f64p]
        L.oipo_load_rrc_csv.restype = C.c_int
        L.oipo_load_rrc_csv.argtypes = [C.c_char_p, C.c_int, f64p]
        L.oipo_remap_cubic_u16.restype = None
        L.oipo_remap_cubic_u16.argtypes = [u16p, C.c_int, C.c_int, C.c_int64, u16p, C.c_int, C.c_int, f32p, f32p]
        L.oipo_cubic_tab.restype = None
        L.oipo_cubic_tab.argtypes = [f32p]
        L.oipo_prestitch_shift.restype = C.c_int64
        L.oipo_prestitch_shift.argtypes = [u16p, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, u16p]
        L.oipo_stitch_concat_u16.restype = None
        L.oipo_stitch_concat_u16.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_int, u16p]
        L.oipo_pan_pipeline.restype = C.c_int64
        L.oipo_pan_pipeline.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p),
                                        f64p, f64p, C.c_int, C.c_int, C.c_int, u16p]
        L.oipo_set_min_process_lines.restype = None
        L.oipo_set_min_process_lines.argtypes = [C.c_int]
        L.oipo_band_align.restype = C.c_int64
        L.oipo_band_align.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int, f64p, f64p, C.c_int, C.c_int64,
                                      C.c_int, C.c_int, u16p]
        L.oipo_stitch_concat_c4.restype = None
        L.oipo_stitch_concat_c4.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_int,
                                            C.c_void_p, u16p]
        L.oipo_unpack_bits.restype = None
        L.oipo_unpack_bits.argtypes = [u8p, C.c_int, C.c_int, C.c_int64, C.c_int64, u16p]
        L.oipo_swap16.restype = None
        L.oipo_swap16.argtypes = [u16p, C.c_int64, u16p]
        L.oipo_mss_split.restype = None
        L.oipo_mss_split.argtypes = [u16p, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        _LIB = L
    return _LIB


def ref_crc_lib():
    """the reference's own CRC.h behind a C shim (oracle/_ref/libref_crc.so), or None."""
    global _REF_CRC
    if _REF_CRC is None:
        so = os.path.join(_HERE, "_ref", "libref_crc.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_crc16_ccitt_false.restype = C.c_uint16
        L.ref_crc16_ccitt_false.argtypes = [u8p, C.c_size_t]
        _REF_CRC = L
    return _REF_CRC


def ref_oip_lib():
    """the reference's own headers compiled behind a C shim (oracle/_ref/libref_oip.so), or None."""
    global _ref_oip
    try:
        return _ref_oip
    except NameError:
        pass
    so = os.path.join(_HERE, "_ref", "libref_oip.so")
    _ref_oip = None
    if os.path.exists(so):
        try:
            L = C.CDLL(so)
            L.ref_inplace_rrc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.ref_inplace_rrc.restype = None
            _ref_oip = L
        except OSError:
            _ref_oip = None
    return _ref_oip


def ref_inplace_rrc(img: np.ndarray, kb: np.ndarray) -> np.ndarray:
    """IMO::InplaceRRC as compiled from the reference's imageop.h (ref imageop.h:129-138); needs oracle/_ref"""
    L = ref_oip_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_oip.so is not built")
    out = np.ascontiguousarray(img, np.uint16).copy()
    kb = np.ascontiguousarray(kb, np.float64)
    L.ref_inplace_rrc(out.ctypes.data_as(C.c_void_p), out.shape[1], out.shape[0], kb.ctypes.data_as(C.c_void_p))
    return out


def _ptr_array(arrs):
    a = (C.c_void_p * len(arrs))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data
    return a


# ---------------------------------------------------------------- convenience wrappers
def crc16(data: np.ndarray) -> int:
    d = np.ascontiguousarray(data, np.uint8)
    return int(lib().oipo_crc16(d, d.size))


def aos_scan(buf: np.ndarray):
    buf = np.ascontiguousarray(buf, np.uint8)
    cap = buf.size // 1024 + 1
    off = np.zeros(cap, np.uint64)
    cnt = np.zeros(3, np.int64)
    n = lib().oipo_aos_scan(buf, buf.size, off, cap, cnt)
    return off[:n].copy(), cnt


def aos_scan_range(buf: np.ndarray, start: int, own_end: int):
    """byte-range shard of the scan: candidates that start in [start, own_end) of buf (which carries the halo after own_end);
    returns (payload offsets, counters, next search position)"""
    buf = np.ascontiguousarray(buf, np.uint8)
    cap = buf.size // 1024 + 1
    off = np.zeros(cap, np.uint64)
    cnt = np.zeros(3, np.int64)
    nxt = C.c_uint64(0)
    n = lib().oipo_aos_scan_range(buf, buf.size, start, own_end, off, cap, cnt, C.byref(nxt))
    return off[:n].copy(), cnt, int(nxt.value)


def imtr_deframe(buf: np.ndarray, payload_off: np.ndarray):
    buf = np.ascontiguousarray(buf, np.uint8)
    payload_off = np.ascontiguousarray(payload_off, np.uint64)
    cap = (payload_off.size * 880 // 882 + 1) * 866
    out = np.zeros(cap, np.uint8)
    st = np.zeros(9, np.int64)
    n = lib().oipo_imtr_deframe(buf, payload_off, payload_off.size, out, cap, st)
    if n < 0:
        raise RuntimeError("oipo_imtr_deframe failed")
    return out[:n].copy(), st


def image_frames(imdt: np.ndarray, tile_cols: int, tile_lines: int):
    imdt = np.ascontiguousarray(imdt, np.uint8)
    g = FrameGeom(tile_cols, tile_lines)
    st = np.zeros(4, np.int64)
    n = lib().oipo_image_frames(imdt, imdt.size, C.byref(g), None, None, None, 0, st)
    if n < 0:
        return n, None, None, None, st
    W = 8 * tile_cols
    aux = np.zeros((n, 48 * 4 * tile_lines), np.uint8)
    pan = np.zeros((n * 4 * tile_lines, W), np.uint16)
    mss = np.zeros((n * tile_lines, W), np.uint16)
    n2 = lib().oipo_image_frames(imdt, imdt.size, C.byref(g), aux.ctypes.data, pan.ctypes.data,
                                 mss.ctypes.data, n, st)
    assert n2 == n
    return n, aux, pan, mss, st


def rrc(img: np.ndarray, kb: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(img, np.uint16).copy()
    h, w = out.shape
    lib().oipo_rrc_u16(out, w, h, np.ascontiguousarray(kb, np.float64).reshape(-1))
    return out


def remap_cubic(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint16)
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    dst = np.zeros((dh, dw), np.uint16)
    lib().oipo_remap_cubic_u16(src, src.shape[1], src.shape[0], src.shape[1], dst, dw, dh, mapx, mapy)
    return dst


def cubic_tab() -> np.ndarray:
    t = np.zeros(128, np.float32)
    lib().oipo_cubic_tab(t)
    return t.reshape(32, 4)


def prestitch_shift(src: np.ndarray, dX: float, dY: float, section_rows=30000, row_guard=32767) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint16)
    h, w = src.shape
    dst = np.zeros_like(src)
    n = lib().oipo_prestitch_shift(src, w, h, dX, dY, section_rows, row_guard, dst)
    if n != h:
        raise RuntimeError(f"oipo_prestitch_shift wrote {n} of {h} rows")
    return dst


def stitch_concat(ccds, fold_half: int) -> np.ndarray:
    ccds = [np.ascontiguousarray(c, np.uint16) for c in ccds]
    h, w = ccds[0].shape
    n = len(ccds)
    out = np.zeros((h, n * w - 2 * (n - 1) * fold_half), np.uint16)
    lib().oipo_stitch_concat_u16(_ptr_array(ccds), n, w, h, fold_half, out)
    return out


def pan_pipeline(ccds, kbs, dX, dY, fold_half, section_rows=30000, row_guard=32767) -> np.ndarray:
    ccds = [np.ascontiguousarray(c, np.uint16) for c in ccds]
    kbs = [np.ascontiguousarray(k, np.float64) for k in kbs]
    h, w = ccds[0].shape
    n = len(ccds)
    out = np.zeros((h, n * w - 2 * (n - 1) * fold_half), np.uint16)
    rc = lib().oipo_pan_pipeline(_ptr_array(ccds), n, w, h, _ptr_array(kbs),
                                 np.ascontiguousarray(dX, np.float64), np.ascontiguousarray(dY, np.float64),
                                 fold_half, section_rows, row_guard, out)
    if rc != h:
        raise RuntimeError(f"oipo_pan_pipeline rc={rc}")
    return out


def band_align(planes, cX, cY, lines_per_section=20000, line_offset=0, overlap=520, keep_leading=False,
               min_process_lines=1500):
    planes = [np.ascontiguousarray(p, np.uint16) for p in planes]
    lines, wb = planes[0].shape
    rows = lines - line_offset - (0 if keep_leading else overlap)
    out = np.zeros((max(rows, 0), wb, 4), np.uint16)
    lib().oipo_set_min_process_lines(min_process_lines)
    n = lib().oipo_band_align(_ptr_array(planes), lines, wb, np.ascontiguousarray(cX, np.float64).reshape(-1),
                              np.ascontiguousarray(cY, np.float64).reshape(-1), lines_per_section, line_offset,
                              overlap, int(keep_leading), out)
    lib().oipo_set_min_process_lines(1500)
    return n, out


def stitch_concat_c4(imgs, fold_half: int, band_map=None) -> np.ndarray:
    imgs = [np.ascontiguousarray(c, np.uint16) for c in imgs]
    h, w, _ = imgs[0].shape
    n = len(imgs)
    out = np.zeros((h, n * w - 2 * (n - 1) * fold_half, 4), np.uint16)
    bm = None
    if band_map is not None:
        bm = (C.c_int * 4)(*band_map)
    lib().oipo_stitch_concat_c4(_ptr_array(imgs), n, w, h, fold_half, bm, out.reshape(-1))
    return out


def unpack_bits(raw: np.ndarray, bits: int, w: int, h: int, pitch: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros((h, w), np.uint16)
    lib().oipo_unpack_bits(raw, bits, w, h, pitch, out)
    return out


def mss_split(mixed: np.ndarray):
    mixed = np.ascontiguousarray(mixed, np.uint16)
    lines, w = mixed.shape
    planes = [np.zeros((lines, w // 4), np.uint16) for _ in range(4)]
    lib().oipo_mss_split(mixed, lines, w, _ptr_array(planes))
    return planes


# ---------------------------------------------------------------------------------------------
# N1: inter-CMOS offset estimation (ref stitcher.h:148-201).  The arithmetic is OpenCV's cv::phaseCorrelate
# (imgproc/src/phasecorr.cpp, version left open by the reference's CMakeLists.txt:8): numpy restatement in fp64,
# pinned against cv2.phaseCorrelate 4.13.0 in tests/test_phasecorr_cpu.py (floating point: tolerance, not bits).
# ---------------------------------------------------------------------------------------------
def optimal_dft_size(n: int) -> int:
    """cv::getOptimalDFTSize: the smallest 2^a 3^b 5^c >= n"""
    best = None
    p2 = 1
    while p2 < 2 * n:
        p3 = p2
        while p3 < 2 * n:
            p5 = p3
            while p5 < 2 * n:
                if p5 >= n and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def phase_correlate(a: np.ndarray, b: np.ndarray):
    """(dx, dy, response) of cv::phaseCorrelate(a, b) without window: zero-pad to the optimal DFT size, normalised
    cross-power spectrum F1 conj(F2) / |F1 conj(F2)|, unscaled inverse DFT, quadrant swap, first maximum,
    5x5 weighted centroid clamped to the image, response = window sum / (M N), result = centre - centroid."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    M, N = optimal_dft_size(a.shape[0]), optimal_dft_size(a.shape[1])
    # odd sizes: OpenCV's fftShift is the circular shift by (M // 2, N // 2) for every parity, which np.roll does below
    pa = np.zeros((M, N)); pa[: a.shape[0], : a.shape[1]] = a
    pb = np.zeros((M, N)); pb[: b.shape[0], : b.shape[1]] = b
    P = np.fft.rfft2(pa) * np.conj(np.fft.rfft2(pb))
    mag = np.abs(P)
    C = np.fft.irfft2(P * mag / (mag * mag + np.finfo(np.float32).eps), s=(M, N)) * (M * N)
    C = np.roll(C, (M // 2, N // 2), axis=(0, 1)).astype(np.float32)
    py, px = np.unravel_index(int(np.argmax(C)), C.shape)
    y0, y1 = max(py - 2, 0), min(py + 2, M - 1)
    x0, x1 = max(px - 2, 0), min(px + 2, N - 1)
    win = C[y0:y1 + 1, x0:x1 + 1].astype(np.float64)
    ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1]
    s = win.sum()
    cx, cy = (xs * win).sum() / (s + np.finfo(np.float64).eps), (ys * win).sum() / (s + np.finfo(np.float64).eps)
    return N / 2.0 - cx, M / 2.0 - cy, s / (M * N)


def stt_parameters(pan1: np.ndarray, pan2: np.ndarray, overlap_cols=200, edge_cols=0, sections=10, lines_per_section=16000,
                   threshold=0.4, max_delta_y=0.0, correlate=phase_correlate):
    """Stitcher::CalcSttParameters (ref stitcher.h:148-201): per-section (line_offset, dx, dy, response, valid) and
    the means over the valid sections (None when there is none: the reference throws)."""
    lines, w = pan1.shape
    gap = (lines - sections * lines_per_section) // (sections + 1)                         # :151
    step = gap + lines_per_section                                                       # :152
    rows = []
    for i in range(sections):
        off = gap + i * step                                                             # :167
        s1 = pan1[off:off + lines_per_section, w - overlap_cols:w - edge_cols].astype(np.float32)   # :175
        s2 = pan2[off:off + lines_per_section, edge_cols:overlap_cols].astype(np.float32)           # :176
        dx, dy, r = correlate(s1, s2)
        ok = r >= threshold and (max_delta_y <= 0.0 or abs(dy) <= max_delta_y)           # :181
        rows.append((off, dx, dy, r, bool(ok)))
    good = [q for q in rows if q[4]]
    mean = None
    if good:
        mean = tuple(sum(q[k] for q in good) / len(good) for k in (1, 2, 3))             # :197-199
    return rows, mean


# ---------------------------------------------------------------------------------------------
# N2: inter-band shift estimation + polynomial fit (ref preproc.h:224-347, :492-550).  cv::resize INTER_CUBIC is
# OpenCV's (restated in resize_cubic_x below, pinned against cv2.resize); the fit is a plain least-squares
# polynomial (NumCpp Poly1d::fit in the reference, numpy.polyfit here), coefficients in ascending order.
# ---------------------------------------------------------------------------------------------
def resize_cubic(src: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """cv::resize(src, Size(cols, rows), 0, 0, INTER_CUBIC) for CV_32FC1: taps floor(f)-1..+2 clamped to the image,
    weights interpolateCubic(A=-0.75) in float, horizontal pass then vertical pass."""
    src = np.asarray(src, np.float32)

    def axis_tab(n_src, n_dst):
        scale = n_src / n_dst
        f = ((np.arange(n_dst) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        x = (f - s).astype(np.float32)
        A = np.float32(-0.75)
        w0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        w1 = ((A + 2) * x - (A + 3)) * x * x + 1
        w2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        w3 = np.float32(1) - w0 - w1 - w2
        idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)
        return idx, np.stack([w0, w1, w2, w3], 1).astype(np.float32)

    ix, wx = axis_tab(src.shape[1], cols)
    iy, wy = axis_tab(src.shape[0], rows)
    h = np.zeros((src.shape[0], cols), np.float32)
    for k in range(4):
        h += src[:, ix[:, k]] * wx[None, :, k]
    out = np.zeros((rows, cols), np.float32)
    for k in range(4):
        out += h[iy[:, k], :] * wy[:, None, k]
    return out


def inter_band_correlation(pan: np.ndarray, bands, slices=10, sections=5, threshold=0.4, corr_lines=16000, min_count=5,
                           correlate=phase_correlate, resize=resize_cubic):
    """PreProcessor::CalcInterBandCorrelation + FilterInterBandShiftValues + DoCorrelationPolynomialFitting.
    Returns (shifts[band][sec*slices+i] = (dx, dy, rs, cx), cX[4][2], cY[4][3]); raises RuntimeError like the
    reference when a band has fewer than min_count usable values (ref :505-510)."""
    lines, w = pan.shape
    base_rows = min(lines, corr_lines)                                                  # :245
    gap = (lines - base_rows * sections) // (sections + 1)                              # :246
    cols = w // slices                                                                  # :247
    brow, bgap, bcols = base_rows // 4, gap // 4, cols // 4                              # :272-274
    shifts = [[None] * (slices * sections) for _ in range(4)]
    for sec in range(sections):
        r0 = gap + sec * (base_rows + gap)                                              # :256
        for i in range(slices):
            base = pan[r0:r0 + base_rows, i * cols:(i + 1) * cols].astype(np.float32)
            for b in range(4):
                q0 = bgap + sec * (brow + bgap)                                         # :283
                sl = bands[b][q0:q0 + brow, i * bcols:(i + 1) * bcols].astype(np.float32)
                up = resize(sl, base_rows, cols)                                        # :299-304
                dx, dy, rs = correlate(base, up)                                        # :314
                shifts[b][sec * slices + i] = (dx, dy, rs, i * cols + cols // 2)        # :322-326
    cX, cY = [], []
    for b in range(4):
        good = [s for s in shifts[b] if s[2] >= threshold]
        if len(good) < min_count:
            raise RuntimeError(f"Not enough valid correlation values for band#{b + 1}: {len(good)} valid values found, "
                               f"{min_count} expected at least")
        x = np.array([s[3] for s in good], np.float64)
        cX.append(np.polyfit(x, np.array([s[0] for s in good]), 1)[::-1].tolist())      # :533 ascending coefficients
        cY.append(np.polyfit(x, np.array([s[1] for s in good]), 2)[::-1].tolist())      # :534
    return shifts, cX, cY


# ---------------------------------------------------------------------------------------------
# Row-window form of the fused PAN path: selected OUTPUT rows of a strip that is too long to run through
# oipo_pan_pipeline as a whole (bench.py's in-run parity check on the 1 M-line strip, BASELINE configs[3]).
# The sections of SectionaryRemap are written out once more here (ref imageop.h:246-272, stitcher.h:83-139, :122-123);
# tests/test_oracle_cpu.py checks this form against the literal whole-strip restatement (oipo_pan_pipeline), which
# is itself pinned against the reference's own compiled PreStitch (tests/golden/ref_prestitch.npz).
# ---------------------------------------------------------------------------------------------
def shift_pieces(total_rows: int, dY: float, section_rows: int = 30000, row_guard: int = 32767):
    """the runs of output rows SectionaryRemap writes, in file order:
    (out_row0, n_rows, dst_row0, sec_off, rows_s, stale_off): output rows [out_row0, +n) = rows [dst_row0, +n) of the
    remapped section buffer; buffer rows [0, rows_s) hold source rows sec_off + t, buffer rows [rows_s, section_rows)
    still hold the previous section's rows stale_off + t (or nothing: stale_off = -1)."""
    if total_rows <= row_guard:                      # the reference throws (imageop.h:242-244); extension: one remap
        return [(0, total_rows, 0, 0, total_rows, -1)], total_rows
    ucut = 0 if dY >= 0.0 else int(-dY) + 1          # stitcher.h:122
    bcut = int(dY) + 1 if dY >= 0.0 else 0           # :123
    cut = ucut + bcut                                # imageop.h:246
    pieces, off, written, prev_off, last = [], 0, 0, -1, None
    s = 0
    while True:
        rows = min(section_rows, total_rows - off)   # :250
        if rows <= cut:                              # :251
            break
        stale = prev_off if rows < section_rows else -1
        if s == 0 and ucut > 0:                      # :260-263
            pieces.append((written, ucut, 0, off, rows, stale))
            written += ucut
        pieces.append((written, rows - cut, ucut, off, rows, stale))     # :265
        written += rows - cut
        last = (off, rows, stale)
        prev_off = off
        off += rows - cut                            # :266
        s += 1
    if bcut > 0 and last is not None:                # :269-272
        pieces.append((written, bcut, section_rows - bcut, last[0], last[1], last[2]))
        written += bcut
    return pieces, section_rows


def pan_rows(gen, n_ccd: int, w: int, kbs, dX, dY, fold_half: int, total_rows: int, rows, section_rows: int = 30000,
             row_guard: int = 32767) -> np.ndarray:
    """output rows `rows` (sorted global line indices) of the fused PAN path of a total_rows-line strip.
    gen(i, a, b) -> raw LE u16 lines [a, b) of CCD i (any lines can be asked for)."""
    rows = np.asarray(rows, np.int64)
    out_w = n_ccd * w - 2 * (n_ccd - 1) * fold_half
    out = np.zeros((rows.size, out_w), np.uint16)
    xo = 0
    for i in range(n_ccd):
        lo, hi = (0 if i == 0 else fold_half), (w if i == n_ccd - 1 else w - fold_half)
        res = np.zeros((rows.size, w), np.uint16)
        # consecutive runs of wanted rows
        cuts = np.flatnonzero(np.diff(rows) != 1) + 1
        runs = np.split(np.arange(rows.size), cuts) if rows.size else []
        if i == 0:                                    # CMOS-1: radiometric correction only
            for idx in runs:
                a, b = int(rows[idx[0]]), int(rows[idx[-1]]) + 1
                res[idx] = rrc(gen(i, a, b), kbs[i])
        else:
            pieces, hbuf = shift_pieces(total_rows, float(dY[i]), section_rows, row_guard)
            m = int(np.ceil(abs(float(dY[i])))) + 4
            for idx in runs:
                g = int(rows[idx[0]])
                g_end = int(rows[idx[-1]]) + 1
                while g < g_end:
                    pc = next(p for p in pieces if p[0] <= g < p[0] + p[1])
                    o0, n, d0, sec_off, rows_s, stale = pc
                    ge = min(g_end, o0 + n)
                    ja, jb = d0 + (g - o0), d0 + (ge - o0)              # rows of the remapped section buffer
                    ta, tb = max(0, ja - m), min(hbuf, jb + m)          # buffer rows the window holds
                    win = np.zeros((tb - ta, w), np.uint16)
                    fa, fb = ta, min(tb, rows_s)                         # fresh rows of this section
                    if fb > fa:
                        win[fa - ta:fb - ta] = rrc(gen(i, sec_off + fa, sec_off + fb), kbs[i])
                    sa = max(ta, rows_s)                                 # rows left over from the previous section
                    if tb > sa and stale >= 0:
                        win[sa - ta:tb - ta] = rrc(gen(i, stale + sa, stale + tb), kbs[i])
                    mx = (np.arange(w)[None, :] + np.zeros((jb - ja, 1)) + float(dX[i])).astype(np.float32)       # stitcher.h:96
                    my = (np.arange(ja, jb)[:, None] + np.zeros((1, w)) + float(dY[i])).astype(np.float32)       # :97
                    my = (my - np.float32(ta)).astype(np.float32)        # exact: an integer off a float with the same ulp
                    sel = idx[(rows[idx] >= g) & (rows[idx] < ge)]
                    res[sel] = remap_cubic(win, mx, my)
                    g = ge
        out[:, xo:xo + hi - lo] = res[:, lo:hi]
        xo += hi - lo
    return out
