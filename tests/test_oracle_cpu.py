"""CPU suite: pins the oracle (oracle/) against everything reference-backed that exists:
  * the reference's own CRC.h compiled stand-alone (oracle/_ref/libref_crc.so) + its check value,
  * cv2.remap / cv2.merge -- the OpenCV calls the reference makes (imageop.h:258, preproc.h:453-464),
    driven by an independent Python restatement of the reference's section loops,
  * committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py),
and checks the host logic / the C-ABI library surface.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from opticalimageprocessor_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2")


# ---------------------------------------------------------------------------------------- CRC
def test_crc_check_value():
    # ref CRC.h:1519 : CRC-16/CCITT-FALSE("123456789") == 0x29B1
    assert oracle.crc16(np.frombuffer(b"123456789", np.uint8)) == 0x29B1
    assert oracle.crc16(np.zeros(0, np.uint8)) == 0xFFFF


def test_crc_vs_reference_crcpp():
    ref = oracle.ref_crc_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_crc.so not built (no /root/reference here)")
    rng = np.random.default_rng(0)
    for n in [0, 1, 2, 7, 876, 890, 1024, 4099]:
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert oracle.crc16(d) == ref.ref_crc16_ccitt_false(d, n)
    d = np.zeros(890, np.uint8)
    assert oracle.crc16(d) == ref.ref_crc16_ccitt_false(d, 890)


def test_crc_golden_and_vectorised_generator():
    g = np.load(os.path.join(GOLD, "crc16.npz"))
    for msg, want in zip(g["msgs"], g["crcs"]):
        assert oracle.crc16(msg) == int(want)
    assert np.array_equal(synth.crc16_rows(g["msgs"]), g["crcs"])


# ---------------------------------------------------------------------------------------- frames
def _small_downlink(seed=3, n_frames=3, tc=16, tl=4, skip=(), **kw):
    imdt, truth = synth.make_imdt(n_frames, tc, tl, seed=seed, skip_seqs=skip)
    imtr = synth.imtr_frames(imdt, chid=0x22)
    aos = synth.aos_frames(imtr.reshape(-1))
    return imdt, truth, imtr, aos


def test_aos_scan_rules():
    imdt, truth, imtr, aos = _small_downlink()
    n = aos.shape[0]
    buf = synth.build_aos_file(aos, empty_every=5, bad_crc_at={2, 7}, bad_inject_at={4}, prefix=b"\x00\x11\x22" * 7)
    off, cnt = oracle.aos_scan(buf)
    n_empty = len(range(0, n, 5))
    assert cnt.tolist() == [n, 3, n_empty]
    pay = np.stack([buf[int(o):int(o) + 880] for o in off])
    assert np.array_equal(pay, aos[:, 14:894])


def test_aos_scan_false_sync_inside_payload_is_shadowed():
    imdt, truth, imtr, aos = _small_downlink()
    aos = aos.copy()
    aos[3, 100:104] = np.frombuffer(synth.AOS_SYNC, np.uint8)  # false sync in a payload
    crc = synth.crc16_rows(aos[3:4, 4:894])
    aos[3, 894], aos[3, 895] = crc[0] >> 8, crc[0] & 0xFF
    buf = aos.reshape(-1)
    off, cnt = oracle.aos_scan(buf)
    assert cnt.tolist() == [aos.shape[0], 0, 0]  # never visited: the scan jumps 1024 bytes past a valid frame
    # the same false sync in a frame that FAILS its CRC is visited and counted invalid
    aos2 = aos.copy()
    aos2[3, 600] ^= 1
    off2, cnt2 = oracle.aos_scan(aos2.reshape(-1))
    assert cnt2.tolist() == [aos.shape[0] - 1, 2, 0]


def test_aos_scan_edges():
    assert oracle.aos_scan(np.zeros(0, np.uint8))[1].tolist() == [0, 0, 0]
    assert oracle.aos_scan(np.zeros(1023, np.uint8))[1].tolist() == [0, 0, 0]
    one = synth.aos_frames(np.arange(880, dtype=np.uint8) % 251)
    assert oracle.aos_scan(one.reshape(-1))[1].tolist() == [1, 0, 0]
    # a sync word in the last 1023 bytes can never be a frame (SURVEY C-5)
    tail = np.concatenate([one.reshape(-1), np.frombuffer(synth.AOS_SYNC + b"\x00" * 500, np.uint8)])
    assert oracle.aos_scan(tail)[1].tolist() == [1, 0, 0]


def test_imtr_and_image_frames_roundtrip():
    tc, tl = 16, 4
    imdt, truth, imtr, aos = _small_downlink(n_frames=4, tc=tc, tl=tl, skip={3})
    buf = synth.build_aos_file(aos, empty_every=7)
    off, cnt = oracle.aos_scan(buf)
    got_imdt, st = oracle.imtr_deframe(buf, off)
    assert st[0] == st[1] and st[2:7].tolist() == [0, 0, 0, 0, 0] and st[7] == 0x22 and st[8] == 1
    assert np.array_equal(got_imdt[:imdt.size], imdt)
    n, aux, pan, mss, fst = oracle.image_frames(got_imdt, tc, tl)
    assert n == 4 and fst.tolist() == [3, 4, 0, 4]
    for s in (1, 2, 4):
        a, p, m = truth[s]
        assert np.array_equal(aux[s - 1], a)
        assert np.array_equal(pan[(s - 1) * 4 * tl:s * 4 * tl], p)
        assert np.array_equal(mss[(s - 1) * tl:s * tl], m)
    assert not aux[2].any() and not pan[2 * 4 * tl:3 * 4 * tl].any() and not mss[2 * tl:3 * tl].any()  # gap -> zeros


def test_imtr_bad_frames_and_restart_rule():
    imdt, truth, imtr, aos = _small_downlink()
    imtr = imtr.copy()
    imtr[1, 0] ^= 0xFF          # bad head signature
    imtr[2, 880] ^= 0xFF        # bad tail signature
    imtr[3, 9] = 0x11           # not image data
    synth.refresh_imtr_crc(imtr[3:4])
    imtr[4, 300] ^= 0x01        # bad CRC
    imtr[6, 4:8] = 0            # seq 0 -> the next accepted frame re-creates the IMDT file
    synth.refresh_imtr_crc(imtr[6:7])
    buf = synth.aos_frames(imtr.reshape(-1)).reshape(-1)
    off, _ = oracle.aos_scan(buf)
    got, st = oracle.imtr_deframe(buf, off)
    nf = imtr.shape[0]
    cut = off.size * 880 // 882
    assert st[0] == cut and st[2:6].tolist() == [1, 1, 1, 1]
    assert st[8] == 2  # first frame + the one after seq 0
    want = imtr[7:cut, 10:876].reshape(-1)
    assert np.array_equal(got, want)


def test_image_frame_incomplete_and_false_signature():
    tc, tl = 16, 4
    imdt, truth = synth.make_imdt(3, tc, tl, seed=9)
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    # drop the head of frame 1 -> "incomplete image frame, ignored" (ref aux_separator.h:289-299)
    cut = imdt[100:]
    n, aux, pan, mss, st = oracle.image_frames(cut, tc, tl)
    assert st.tolist() == [2, 3, 1, 3] and n == 3  # frame 1 missing -> zero filled as a gap
    assert not pan[:4 * tl].any()
    # a trailer signature inside pixel data makes that frame AND the hunt position skip (sequential memmem)
    bad = imdt.copy()
    pos = frame_bytes + 192 * tl + 64
    bad[pos:pos + 4] = np.frombuffer(synth.IMG_SIG, np.uint8)
    n2, *_, st2 = oracle.image_frames(bad, tc, tl)
    assert st2[2] >= 1  # at least the false hit is reported incomplete


# ---------------------------------------------------------------------------------------- RRC
def test_rrc_semantics():
    img = np.array([[1, 0, 40000, 7, 100, 65535]], np.uint16)
    kb = np.array([[1.0, -0.5], [1.0, -1.0], [2.0, 0.0], [1.0, 4294967296.0], [0.999, 0.9], [1.0, 0.0]])
    # trunc toward zero; negative wraps; >65535 wraps; beyond int32 -> 0
    assert oracle.rrc(img, kb).tolist() == [[0, 65535, 14464, 0, 100, 65535]]


def test_rrc_csv_roundtrip(tmp_path):
    kb = synth.rrc_coeffs(64, 5)
    p = str(tmp_path / "PAN-1.csv")
    synth.write_rrc_csv(p, kb)
    got = np.zeros((64, 2))
    assert oracle.lib().oipo_load_rrc_csv(p.encode(), 64, got.reshape(-1)) == 0
    assert np.array_equal(got, kb)
    assert oracle.lib().oipo_load_rrc_csv(p.encode(), 63, got.reshape(-1)) == -3


# ---------------------------------------------------------------------------------------- remap
def _cv_remap(src, mx, my):
    return cv2.remap(src, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)


@pytest.mark.parametrize("dX,dY", [(0, 0), (3, -2), (0.5, 0.5), (1.37, -2.61), (-1.984375, 4.015625), (0.015625, 0),
                                   (-5.3, 7.77), (700.2, 0.1), (0.2, -300.9)])
def test_remap_bitexact_vs_cv2(dX, dY):
    rng = np.random.default_rng(1)
    for (W, H) in [(640, 257), (12288, 24), (5, 5), (3, 9)]:
        src = rng.integers(0, 65536, (H, W), dtype=np.uint16)
        mx = (np.arange(W)[None, :] + np.zeros((H, 1)) + dX).astype(np.float32)
        my = (np.arange(H)[:, None] + np.zeros((1, W)) + dY).astype(np.float32)
        assert np.array_equal(oracle.remap_cubic(src, mx, my), _cv_remap(src, mx, my))


def test_remap_random_maps_vs_cv2():
    rng = np.random.default_rng(2)
    src = rng.integers(0, 65536, (97, 211), dtype=np.uint16)
    mx = rng.uniform(-6, 217, (64, 300)).astype(np.float32)
    my = rng.uniform(-6, 103, (64, 300)).astype(np.float32)
    assert np.array_equal(oracle.remap_cubic(src, mx, my), _cv_remap(src, mx, my))


def test_cubic_weight_table_vs_cv2():
    """cv2.remap on a float impulse image returns OpenCV's own 2-D weight table entries exactly"""
    tab = oracle.cubic_tab()
    src = np.zeros((9, 9), np.float32)
    src[4, 4] = 1.0
    for fx in range(32):
        for fy in (0, 7, 16, 31):
            mx = np.full((1, 1), 4 + fx / 32.0, np.float32)
            my = np.full((1, 1), 4 + fy / 32.0, np.float32)
            got = cv2.remap(src, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)[0, 0]
            assert got == np.float32(tab[fy, 1]) * np.float32(tab[fx, 1])
    g = np.load(os.path.join(GOLD, "cubic_tab.npz"))
    assert np.array_equal(tab, g["tab"])


def test_remap_golden_fixture():
    g = np.load(os.path.join(GOLD, "remap_cv2.npz"))
    src = g["src"]
    H, W = src.shape
    for i, (dX, dY) in enumerate(g["shifts"]):
        mx = (np.arange(W)[None, :] + np.zeros((H, 1)) + dX).astype(np.float32)
        my = (np.arange(H)[:, None] + np.zeros((1, W)) + dY).astype(np.float32)
        assert np.array_equal(oracle.remap_cubic(src, mx, my), g["out"][i])


# ------------------------------------------------- reference section loops restated over cv2.remap
def _prestitch_cv2(src, dX, dY, S, G):
    """Stitcher::PreStitch + IMO::SectionaryRemap restated in Python on top of the real cv2.remap
    (ref stitcher.h:83-139, imageop.h:230-275), written independently of oracle/oip_oracle.c"""
    T, W = src.shape
    if T <= G:
        mx = (np.arange(W)[None, :] + np.zeros((T, 1)) + dX).astype(np.float32)
        my = (np.arange(T)[:, None] + np.zeros((1, W)) + dY).astype(np.float32)
        return _cv_remap(src, mx, my)
    buff = np.zeros((S, W), np.uint16)
    mx = (np.arange(W)[None, :] + np.zeros((S, 1)) + dX).astype(np.float32)
    my = (np.arange(S)[:, None] + np.zeros((1, W)) + dY).astype(np.float32)
    ucut = 0 if dY >= 0 else int(-dY) + 1
    bcut = int(dY) + 1 if dY >= 0 else 0
    cut = ucut + bcut
    out = []
    off, s, dst = 0, 0, None
    while True:
        rows = min(S, T - off)
        if rows <= cut:
            break
        buff[:rows] = src[off:off + rows]
        dst = _cv_remap(buff, mx, my)
        if s == 0 and ucut > 0:
            out.append(dst[:ucut])
        out.append(dst[ucut:rows - bcut])
        off += rows - cut
        s += 1
    if bcut > 0:
        out.append(dst[S - bcut:S])
    return np.concatenate(out)


@pytest.mark.parametrize("dX,dY", [(1.37, -2.61), (-0.83, 3.19), (0.0, 0.0), (2.5, 40.25), (-3.0, -17.5)])
@pytest.mark.parametrize("rows", [1500, 1337, 953, 449])
def test_sectioned_shift_vs_cv2_loop(dX, dY, rows):
    rng = np.random.default_rng(11)
    src = rng.integers(0, 65536, (rows, 96), dtype=np.uint16)
    want = _prestitch_cv2(src, dX, dY, 400, 450)
    assert want.shape == src.shape
    assert np.array_equal(oracle.prestitch_shift(src, dX, dY, 400, 450), want)


def test_sectioned_shift_reference_geometry():
    """real section size (30000 rows, guard 32767) on a narrow strip: 2 sections + stale bottom rows"""
    rng = np.random.default_rng(12)
    src = rng.integers(0, 4096, (32768, 16), dtype=np.uint16)
    for dX, dY in [(1.37, -2.61), (-0.83, 3.19)]:
        assert np.array_equal(oracle.prestitch_shift(src, dX, dY), _prestitch_cv2(src, dX, dY, 30000, 32767))


def _band_align_cv2(planes, cX, cY, lps, line_offset, overlap, keep, min_lines):
    """PreProcessor::DoInterBandAlignment (both overloads) over cv2.remap + cv2.merge (ref preproc.h:351-468)"""
    lines, wb = planes[0].shape
    out = np.zeros((lines - line_offset - (0 if keep else overlap), wb, 4), np.uint16)
    offset, processed, i = line_offset, 0, 0
    while True:
        n = min(lines - offset, lps)
        if lines < offset or n < min_lines:
            break
        bands = []
        for b in range(4):
            xx = (np.arange(wb, dtype=np.int64) * 4)[None, :].astype(np.float64)
            yy = (np.arange(n, dtype=np.int64) * 4)[:, None].astype(np.float64)
            mx = ((cX[b][1] * xx + cX[b][0] + xx) / 4 + 0 * yy).astype(np.float32)
            my = ((cY[b][2] * xx * xx + cY[b][1] * xx + cY[b][0] + yy) / 4).astype(np.float32)
            bands.append(_cv_remap(np.ascontiguousarray(planes[b][offset:offset + n]), mx, my))
        sec = cv2.merge(bands)
        if i == 0 and keep:
            out[:overlap] = sec[:overlap]
            processed += overlap
        out[processed:processed + n - overlap] = sec[overlap:n]
        processed += n - overlap
        offset += lps - overlap
        i += 1
    return processed, out


@pytest.mark.parametrize("keep", [False, True])
@pytest.mark.parametrize("lines,lps,overlap,off", [(700, 300, 40, 0), (650, 256, 32, 10), (300, 400, 20, 0)])
def test_band_align_vs_cv2_loop(keep, lines, lps, overlap, off):
    rng = np.random.default_rng(21)
    wb = 96
    planes = [rng.integers(0, 65536, (lines, wb), dtype=np.uint16) for _ in range(4)]
    cX = [[0.8 + 0.1 * b, -1.5e-3 * (b + 1)] for b in range(4)]
    cY = [[-3.2 + b, 2e-3 * (b + 1), -1e-5 * (b - 1.5)] for b in range(4)]
    n_want, want = _band_align_cv2(planes, cX, cY, lps, off, overlap, keep, 64)
    n_got, got = oracle.band_align(planes, cX, cY, lps, off, overlap, keep, min_process_lines=64)
    assert n_got == n_want
    assert np.array_equal(got[:n_got], want[:n_want])


def test_band_align_argument_errors():
    planes = [np.zeros((2000, 8), np.uint16) for _ in range(4)]
    z2, z3 = np.zeros((4, 2)), np.zeros((4, 3))
    assert oracle.band_align(planes, z2, z3, overlap=3001)[0] == -1      # ref preproc.h:355
    assert oracle.band_align(planes, z2, z3, lines_per_section=32768)[0] == -2  # :359
    assert oracle.band_align(planes, z2, z3, lines_per_section=1000)[0] == -3   # :362
    assert oracle.band_align(planes, z2, z3, line_offset=600)[0] == -4          # :365


def test_concat_and_mss_split():
    rng = np.random.default_rng(4)
    a, b = rng.integers(0, 65536, (5, 40), dtype=np.uint16), rng.integers(0, 65536, (5, 40), dtype=np.uint16)
    out = oracle.stitch_concat([a, b], 3)
    assert np.array_equal(out, np.concatenate([a[:, :37], b[:, 3:]], axis=1))  # ref imageop.h:340-350
    c = rng.integers(0, 65536, (5, 40), dtype=np.uint16)
    out3 = oracle.stitch_concat([a, b, c], 3)
    assert np.array_equal(out3, np.concatenate([a[:, :37], b[:, 3:37], c[:, 3:]], axis=1))
    m = rng.integers(0, 65536, (6, 32), dtype=np.uint16)
    pl = oracle.mss_split(m)
    for k in range(4):
        assert np.array_equal(pl[k], m[:, 8 * k:8 * k + 8])  # ref preproc.h:69-75
    i1, i2 = rng.integers(0, 65536, (4, 10, 4), dtype=np.uint16), rng.integers(0, 65536, (4, 10, 4), dtype=np.uint16)
    o = oracle.stitch_concat_c4([i1, i2], 2, band_map=[3, 2, 1, 4])
    want = np.concatenate([i1[:, :8], i2[:, 2:]], axis=1)[:, :, [2, 1, 0, 3]]
    assert np.array_equal(o, want)


def test_unpack_bits_extension():
    rng = np.random.default_rng(6)
    for bits in (10, 12):
        img = rng.integers(0, 1 << bits, (7, 64), dtype=np.uint16)
        raw = synth.pack_bits(img, bits)
        assert np.array_equal(oracle.unpack_bits(raw, bits, 64, 7, raw.shape[1]), img)


# ---------------------------------------------------------------------------------------- C ABI surface
def test_capi_library_exports_every_declared_symbol():
    from opticalimageprocessor_b200 import build, capi
    build.build()
    L = capi.load()  # binds every name in capi.SYMBOLS or raises
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "oip_b200.h")).read()
    declared = set(re.findall(r"\b(oip_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"oip_status"}
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/oip_b200.h but not exported"
        assert name in capi.SYMBOLS, f"{name} has no ctypes prototype"
    assert L.oip_abi_version() == 2


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the drop-in boundary is a C ABI: include/oip_b200.h compiles as pedantic C99 (no C++ or torch types in a signature) and
    a C program links against the in-tree library -- the host-only entry points run without a GPU, oip_ctx_create says why
    it cannot"""
    import subprocess
    import torch
    from opticalimageprocessor_b200 import build
    build.build()
    root = os.path.join(os.path.dirname(__file__), "..")
    src = tmp_path / "abi.c"
    src.write_text("""
#include <stdio.h>
#include <string.h>
#include "oip_b200.h"
int main(void) {
    oip_ctx *ctx = NULL;
    oip_frame_geom g = {16, 4};
    int64_t st[4] = {-1, -1, -1, -1};
    if (oip_abi_version() != 2) return 10;
    if (oip_pan_out_width(3, 8192, 100) != 3 * 8192 - 4 * 100) return 11;
    if (oip_image_frames_chain(NULL, NULL, 0, 1000000, &g, NULL, 0, st) != OIP_OK || st[1] != 0) return 12;
    if (oip_ctx_create(0, NULL, 1, &ctx) == OIP_OK) { oip_ctx_destroy(ctx); puts("gpu"); return 0; }
    if (!strstr(oip_last_error(), "no CPU fallback")) return 13;
    puts("no gpu");
    return 0;
}
""")
    exe = str(tmp_path / "abi")
    lib_dir = os.path.dirname(build.LIB)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", exe,
                           "-L" + lib_dir, "-loip_b200", "-Wl,-rpath," + lib_dir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.strip() == ("gpu" if torch.cuda.is_available() else "no gpu")


def test_capi_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from opticalimageprocessor_b200 import capi
    L = capi.load()
    h = C.c_void_p()
    rc = L.oip_ctx_create(0, None, 1, C.byref(h))
    assert rc == capi.OIP_E_CUDA and b"no CPU fallback" in L.oip_last_error()


def test_pan_rows_needed_planner():
    from opticalimageprocessor_b200 import capi
    L = capi.load()
    d = capi.PanDesc()
    d.n_ccd, d.w, d.total_rows, d.row0, d.n_rows = 2, 64, 70000, 35000, 100
    d.fold_half, d.section_rows, d.row_guard = 4, 30000, 32767
    for i in range(2):
        d.ccd[i].fmt, d.ccd[i].n_seg = 0, 1
    d.ccd[1].shifted, d.ccd[1].dX, d.ccd[1].dY = 1, 1.37, -2.61
    f, l, sf, sl = (C.c_int64() for _ in range(4))
    assert L.oip_pan_rows_needed(C.byref(d), 0, f, l, sf, sl) == 0
    assert (f.value, l.value) == (35000, 35100)
    assert L.oip_pan_rows_needed(C.byref(d), 1, f, l, sf, sl) == 0
    # dY=-2.61: taps start at floor(y-2.61)-1 = y-4 and end at y-3+2
    assert (f.value, l.value) == (35000 - 4, 35099 - 3 + 2 + 1)
    assert (sf.value, sl.value) == (0, 0)


# ------------------------------------------------- row-window form used by bench.py's in-run parity check
@pytest.mark.parametrize("total", [1500, 1337, 953, 449])
@pytest.mark.parametrize("dX,dY", [([0, 1.37, -0.83], [0, -2.61, 3.19]), ([0, 2.5, -3.0], [0, 40.25, -17.5])])
def test_pan_rows_window_form_equals_whole_strip(total, dX, dY):
    """oracle.pan_rows (selected output rows from source-row windows: section edges, stale rows of a partial last section)
    == the literal whole-strip restatement oipo_pan_pipeline (pinned against the reference's compiled PreStitch)"""
    from opticalimageprocessor_b200 import synth
    n, w, f, S, G = 3, 96, 10, 400, 450
    ccds = [np.random.default_rng(5 + i).integers(0, 65536, (total, w), dtype=np.uint16) for i in range(n)]
    kbs = [synth.rrc_coeffs(w, 70 + i) * np.array([1 / 16.0, 1.0]) for i in range(n)]
    want = oracle.pan_pipeline(ccds, kbs, dX, dY, f, S, G)
    gen = lambda i, a, b: ccds[i][a:b]
    assert np.array_equal(oracle.pan_rows(gen, n, w, kbs, dX, dY, f, total, np.arange(total), S, G), want)
    sub = np.unique(np.concatenate([np.arange(16), np.arange(total - 16, total), np.random.default_rng(1).integers(0, total, 60)]))
    assert np.array_equal(oracle.pan_rows(gen, n, w, kbs, dX, dY, f, total, sub, S, G), want[sub])


def test_pan_rows_reference_geometry_section_edges():
    """30000-row sections / guard 32767 on a narrow 32768-line strip: the rows around the section edge and the stale
    bottom rows, window form == whole strip"""
    from opticalimageprocessor_b200 import synth
    n, w, f, total = 2, 32, 4, 32768
    ccds = [synth.strip_dn(w, total, 40 + i) for i in range(n)]
    kbs = [synth.rrc_coeffs(w, 50 + i) for i in range(n)]
    for dX, dY in [([0, 1.37], [0, -2.61]), ([0, -0.83], [0, 3.19])]:
        want = oracle.pan_pipeline(ccds, kbs, dX, dY, f)
        rows = np.unique(np.concatenate([np.arange(24), np.arange(29980, 30020), np.arange(total - 24, total)]))
        got = oracle.pan_rows(lambda i, a, b: ccds[i][a:b], n, w, kbs, dX, dY, f, total, rows)
        assert np.array_equal(got, want[rows])
