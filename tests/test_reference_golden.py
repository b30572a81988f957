"""Parity against outputs of THE REFERENCE ITSELF, run in the build container through
oracle/_ref/libref_oip.so (reference headers compiled unmodified; tests/golden/make_golden_ref.py).

CPU:  the oracle must reproduce the reference's files byte for byte (sha256 fixtures).
GPU:  the CUDA path must reproduce the same fixtures (marked gpu)."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

import oracle
from opticalimageprocessor_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _inputs():
    """input generators shared with the fixture script (without loading the reference library)"""
    src = open(os.path.join(GOLD, "make_golden_ref.py")).read()
    ns = {}
    # only the two pure generator functions are needed
    import textwrap
    code = "import numpy as np\nfrom opticalimageprocessor_b200 import synth\nW = 12288\n"
    for fn in ("auxsep_input", "prestitch_input"):
        a = src.index(f"def {fn}(")
        b = src.index("\n\n\n", a)
        code += src[a:b] + "\n\n"
    exec(compile(code, "golden_inputs", "exec"), ns)
    return ns


@pytest.fixture(scope="module")
def auxsep_case():
    g = np.load(os.path.join(GOLD, "ref_auxsep.npz"))
    buf = _inputs()["auxsep_input"]()
    assert buf.size == int(g["input_bytes"]) and sha(buf) == str(g["input_sha256"]), "synthetic input drifted"
    return g, buf


def test_oracle_reproduces_reference_auxsep(auxsep_case):
    """AuxSeparator::Separate() (ref aux_separator.h:224-245) ran on this exact downlink; the oracle's
    three stages must give the same IMDT / AUX / PAN.RAW / MSS.RAW bytes"""
    g, buf = auxsep_case
    off, cnt = oracle.aos_scan(buf)
    imdt, st = oracle.imtr_deframe(buf, off)
    assert imdt.size == int(g["imdt_bytes"]) and sha(imdt) == str(g["imdt_sha256"])
    assert ("CMOS-1" if st[7] == 0x11 else "CMOS-2") in str(g["imdt_name"])       # ref aux_separator.h:513-523
    n, aux, pan, mss, fst = oracle.image_frames(imdt, 1536, 256)
    assert aux.size == int(g["aux_bytes"]) and sha(aux) == str(g["aux_sha256"])
    assert pan.size * 2 == int(g["pan_bytes"]) and sha(pan) == str(g["pan_sha256"])
    assert mss.size * 2 == int(g["mss_bytes"]) and sha(mss) == str(g["mss_sha256"])


@pytest.mark.gpu
def test_gpu_reproduces_reference_auxsep(ctx, auxsep_case):
    import torch
    from opticalimageprocessor_b200 import ops
    g, buf = auxsep_case
    d = torch.from_numpy(buf).cuda()
    off, cnt = ops.aos_scan(ctx, d)
    imdt, st = ops.imtr_deframe(ctx, d, off)
    assert sha(imdt.cpu().numpy()) == str(g["imdt_sha256"])
    ents, fst = ops.image_frames_index(ctx, imdt, 1536, 256)
    aux, pan, mss = ops.unpack_frames(ctx, imdt, 1536, 256, ents, int(fst[1]))
    ctx.sync()
    assert sha(aux.cpu().numpy()) == str(g["aux_sha256"])
    assert sha(pan.cpu().numpy()) == str(g["pan_sha256"])
    assert sha(mss.cpu().numpy()) == str(g["mss_sha256"])


def _check_prestitch(out, g, tag):
    rows = int(g["rows"])
    blocks = [sha(out[i:i + 1024]) for i in range(0, rows, 1024)]
    bad = [i for i, (a, b) in enumerate(zip(blocks, g[tag + "_block_sha"])) if a != str(b)]
    idx = g[tag + "_rows_idx"]
    rows_bad = [int(r) for r, want in zip(idx, g[tag + "_rows"]) if not np.array_equal(out[r], want)]
    assert not bad and not rows_bad, f"{tag}: blocks {bad} / rows {rows_bad} differ from the reference run"


@pytest.mark.parametrize("tag", ["neg", "pos"])
def test_oracle_reproduces_reference_prestitch(tag):
    """Stitcher::PreStitch + SectionaryRemap (ref stitcher.h:83-139, imageop.h:230-275) at the reference's
    real geometry: 12288 px, 30000-row sections, 32768 lines -> section edge + stale bottom rows"""
    p = os.path.join(GOLD, "ref_prestitch.npz")
    if not os.path.exists(p):
        pytest.skip("ref_prestitch.npz not generated")
    g = np.load(p)
    src = _inputs()["prestitch_input"](int(g["rows"]), int(g["seed"]))
    dx, dy = g[tag + "_shift"]
    _check_prestitch(oracle.prestitch_shift(src, float(dx), float(dy)), g, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["neg", "pos"])
def test_gpu_reproduces_reference_prestitch(ctx, tag):
    import torch
    from opticalimageprocessor_b200 import ops
    p = os.path.join(GOLD, "ref_prestitch.npz")
    if not os.path.exists(p):
        pytest.skip("ref_prestitch.npz not generated")
    g = np.load(p)
    src = _inputs()["prestitch_input"](int(g["rows"]), int(g["seed"]))
    dx, dy = g[tag + "_shift"]
    out = ops.prestitch_shift(ctx, torch.from_numpy(src).cuda(), float(dx), float(dy)).cpu().numpy()
    _check_prestitch(out, g, tag)


def test_reference_rrc_direct():
    """IMO::InplaceRRC itself (ref imageop.h:129-138), when the checker library is present"""
    so = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_oip.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_oip.so not built (no /root/reference here)")
    import ctypes as C
    L = C.CDLL(so)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 65536, (64, 12288), dtype=np.uint16)
    kb = synth.rrc_coeffs(12288, 9)
    kb[3] = (1.0, -0.5)
    kb[4] = (2.0, 0.0)       # wraps past 65535 exactly like the reference's x86 build
    want = img.copy()
    L.ref_inplace_rrc(want.ctypes.data_as(C.c_void_p), 12288, 64, kb.ctypes.data_as(C.c_void_p))
    assert np.array_equal(oracle.rrc(img, kb), want)


def test_reference_rrc_direct_extreme_coefficients():
    """the same comparison over coefficient extremes: zero / negative / denormal / huge gains, biases at and beyond the
    uint16 and int32 ranges, NaN and infinities -- the oracle restates what the reference's compiled (double)->(uint16)
    conversion does in every one of these cases (ref imageop.h:134-136)"""
    so = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_oip.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_oip.so not built (no /root/reference here)")
    import ctypes as C
    L = C.CDLL(so)
    rng = np.random.default_rng(1)
    W = 12288
    for it in range(10):
        img = rng.integers(0, 65536, (8, W), dtype=np.uint16)
        k, b = rng.normal(1.0, 0.2, W), rng.normal(0, 50, W)
        idx = rng.choice(W, 600, replace=False)
        k[idx] = rng.choice([0.0, -1.0, -0.0, 1e-300, 1e5, -1e5, 32768.5, 65536.0, 1e10, -1e10, np.nan, np.inf, -np.inf, 4.9e-324], 600)
        b[idx[::2]] = rng.choice([0.0, -0.5, 0.5, 65535.0, 65536.0, -65536.0, 2147483647.0, 2147483648.0, -2147483649.0, 1e19, -1e19,
                                  np.nan, np.inf, -np.inf], 300)
        kb = np.ascontiguousarray(np.stack([k, b], 1))
        want = img.copy()
        L.ref_inplace_rrc(want.ctypes.data_as(C.c_void_p), W, 8, kb.ctypes.data_as(C.c_void_p))
        assert np.array_equal(oracle.rrc(img, kb), want), it


# ------------------------------------------------------------------------------------------------
# band alignment and the RRC CSV loader against the reference's own preproc.h / imageop.h (round 2):
# tests/golden/ref_bandalign.npz and ref_rrc_csv.npz were written by PreProcessor::LoadMSS + DoRRC4MSS +
# DoInterBandAlignment and IMO::LoadRRCParamFile compiled unmodified under oracle/_ref
# ------------------------------------------------------------------------------------------------
def _golden_consts():
    src = open(os.path.join(GOLD, "make_golden_ref.py")).read()
    ns = {}
    a = src.index("BA_CX = ")
    b = src.index("def make_bandalign")
    c = src.index("RRC_CSV_TEXTS = ")
    d = src.index("def make_rrccsv")
    exec(compile("import numpy as np\n" + src[a:b] + src[c:d], "golden_consts", "exec"), ns)
    return ns


def _check_bandalign(out, g, tag):
    rows = int(g[tag + "_rows"])
    assert out.shape == (rows, 3072, 4)
    blocks = [sha(out[i:i + 512]) for i in range(0, rows, 512)]
    bad = [i for i, (a, b) in enumerate(zip(blocks, g[tag + "_block_sha"])) if a != str(b)]
    assert not bad, f"{tag}: 512-row blocks {bad} differ from the reference's DoInterBandAlignment run"


BA_TAGS = ["s3000", "s3000_keep_off_norrc", "default"]


@pytest.mark.parametrize("tag", BA_TAGS)
def test_oracle_reproduces_reference_band_alignment(tag):
    """ref preproc.h:56-80 (band split), :202-222 (per-band RRC), :351-468 (sections, polynomial map, remap, merge)"""
    g = np.load(os.path.join(GOLD, "ref_bandalign.npz"))
    ns = _golden_consts()
    lines, lps, off, ov, keep, do_rrc, seed = (int(v) for v in g[tag + "_params"])
    mss = ns["bandalign_input"](lines, seed)
    planes = oracle.mss_split(mss)
    if do_rrc:
        planes = [oracle.rrc(p, synth.rrc_coeffs(3072, 300 + b)) for b, p in enumerate(planes)]
    n, out = oracle.band_align(planes, g["cX"], g["cY"], lines_per_section=lps, line_offset=off, overlap=ov, keep_leading=bool(keep))
    _check_bandalign(out, g, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", BA_TAGS)
@pytest.mark.parametrize("fast", [1, 0])
def test_gpu_reproduces_reference_band_alignment(ctx, tag, fast):
    import torch
    from opticalimageprocessor_b200 import ops
    g = np.load(os.path.join(GOLD, "ref_bandalign.npz"))
    ns = _golden_consts()
    lines, lps, off, ov, keep, do_rrc, seed = (int(v) for v in g[tag + "_params"])
    mss = torch.from_numpy(ns["bandalign_input"](lines, seed)).cuda()
    kbs = [torch.from_numpy(synth.rrc_coeffs(3072, 300 + b)).cuda() for b in range(4)] if do_rrc else None
    ctx.set_option("mss_fast", fast)
    try:
        n, out = ops.band_align(ctx, mss, 3072, kbs, g["cX"], g["cY"], lines_per_section=lps, line_offset=off, overlap=ov,
                                keep_leading=bool(keep))
        ctx.sync()
    finally:
        ctx.set_option("mss_fast", 1)
    _check_bandalign(out.cpu().numpy(), g, tag)


def test_rrc_csv_loaders_match_reference(tmp_path):
    """IMO::LoadRRCParamFile (ref imageop.h:140-192) ran on these texts: the oracle's and the product's host parser
    (oip_load_rrc_csv, no GPU involved) must accept / reject the same files and return the same doubles"""
    import ctypes as C
    from opticalimageprocessor_b200 import capi
    g = np.load(os.path.join(GOLD, "ref_rrc_csv.npz"))
    ns = _golden_consts()
    L = capi.load()
    for tag, text in ns["RRC_CSV_TEXTS"].items():
        p = tmp_path / (tag + ".csv")
        p.write_bytes(text.encode())
        n = ns["RRC_CSV_EXPECT"][tag]
        want_ok = bool(g[tag + "_ok"])
        kb_o = np.zeros(2 * n)
        rc_o = oracle.lib().oipo_load_rrc_csv(str(p).encode(), n, kb_o)
        kb_p = np.zeros(2 * n)
        rc_p = L.oip_load_rrc_csv(str(p).encode(), n, kb_p.ctypes.data)
        assert (rc_o == 0) == want_ok, (tag, rc_o)
        assert (rc_p == 0) == want_ok, (tag, rc_p, capi.last_error())
        if want_ok:
            assert np.array_equal(kb_o, g[tag + "_kb"]) and np.array_equal(kb_p, g[tag + "_kb"]), tag
    assert not bool(g["missing_file_ok"])
    assert L.oip_load_rrc_csv(str(tmp_path / "does_not_exist.csv").encode(), 2, np.zeros(4).ctypes.data) == capi.OIP_E_IO


def test_rrc_csv_loader_fuzz_against_the_reference_parser(tmp_path):
    """1500 random CSV texts -- CRLF / blank lines, odd separators, leading and trailing blanks, numbers in every spelling
    strtod knows and a few it does not, wrong headers, too few rows -- go through IMO::LoadRRCParamFile (compiled unmodified,
    oracle/_ref) and through oip_load_rrc_csv: both accept the same texts and return the same doubles.  Texts with MORE rows
    than columns are left out: the reference writes them past the end of its array (glibc aborts with heap corruption);
    oip_load_rrc_csv refuses them with the reference's own message and touches nothing beyond the caller's buffer."""
    so = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_oip.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_oip.so not built (no /root/reference here)")
    import ctypes as C
    from opticalimageprocessor_b200 import capi
    REF = C.CDLL(so)
    REF.ref_load_rrc_csv.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    lib = capi.load()
    rng = np.random.default_rng(0)
    nums = ["1", "0", "-1", "1.5", "0.987654321012", "1e0", "1E1", "-3e-3", ".5", "5.", "+2", "1e400", "-1e-400", "nan", "inf", "0x10", "1,5",
            "", " ", "abc", "1.2.3", "1e", "--1"]
    seps = [" , ", ",", " ,", ", ", "\t,\t", " ", ";", " , , ", ",,"]
    eols = ["\n", "\r\n", "\n\n", "\r"]
    p = str(tmp_path / "t.csv")
    accepted = 0
    for it in range(1500):
        n = int(rng.integers(1, 6))
        eol = eols[int(rng.choice(len(eols), p=[0.6, 0.25, 0.1, 0.05]))]
        head = [str(rng.choice(["1", "2", "x", ""], p=[0.85, 0.05, 0.05, 0.05])),
                str(rng.choice([str(n), str(n + 1), "0", "n", ""], p=[0.8, 0.05, 0.05, 0.05, 0.05])),
                str(rng.choice(["0", "1", "", "zero"], p=[0.85, 0.05, 0.05, 0.05]))]
        if rng.random() < 0.05:
            head = head[:int(rng.integers(0, 3))]
        rows = []
        for r in range(int(rng.choice([n, n - 1, 0], p=[0.9, 0.07, 0.03]))):
            good = rng.random() < 0.93
            a = str(rng.choice(nums[:11])) if good else str(rng.choice(nums))
            b = str(rng.choice(nums[:11])) if good else str(rng.choice(nums))
            sep = seps[0] if good and rng.random() < 0.7 else str(rng.choice(seps))
            tail = "" if rng.random() < 0.9 else str(rng.choice([" extra", " ,3", "   ", "\t"]))
            lead = "" if rng.random() < 0.9 else str(rng.choice(["  ", "\t"]))
            rows.append(lead + a + sep + b + tail)
        text = eol.join(head + rows) + (eol if rng.random() < 0.8 else "")
        with open(p, "wb") as f:
            f.write(text.encode())
        kr, kg = (C.c_double * (2 * n))(), (C.c_double * (2 * n))()
        rc_r = REF.ref_load_rrc_csv(p.encode(), n, kr)
        rc_g = lib.oip_load_rrc_csv(p.encode(), n, kg)
        assert (rc_r == 0) == (rc_g == 0), (it, rc_r, rc_g, text)
        if rc_r == 0:
            accepted += 1
            assert np.array_equal(np.array(list(kr)), np.array(list(kg)), equal_nan=True), (it, text, list(kr), list(kg))
    assert 300 < accepted < 1300, accepted
    # more rows than columns: refused, nothing written past the two expected pairs
    with open(p, "w") as f:
        f.write("1\n2\n0\n1 , 0\n2 , 1\n3 , 2\n4 , 3\n")
    kg = (C.c_double * 6)(*([-7.0] * 6))
    assert lib.oip_load_rrc_csv(p.encode(), 2, kg) == capi.OIP_E_INVALID and list(kg)[4:] == [-7.0, -7.0]
    assert b"2 lines of param expected, 4 lines parsed" in lib.oip_last_error()


def test_reference_imdt_after_a_mid_stream_restart(tmp_path):
    """A frame with sequence number 0 makes the reference re-create the IMDT file (ref aux_separator.h:513-528:
    `imdt.attach(fopen(name, "wb"))`).  The new stream truncates the file, but the OLD stream is only flushed afterwards, at its
    old offset: when more data preceded the restart than follows it, the reference's file is [data after the restart] + a
    hole of zeros + the last buffered bytes of the data before it.  The oracle (and liboip) write the data after the last
    restart and nothing else -- the documented deviation; the two agree byte for byte whenever the data
    after the restart is the longer part, and the reference's file always STARTS with the oracle's."""
    so = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_oip.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_oip.so not built (no /root/reference here)")
    import ctypes as C
    REF = C.CDLL(so)
    REF.ref_auxsep.argtypes = [C.c_char_p, C.c_char_p]
    rng = np.random.default_rng(3)
    n_imtr = 3000
    payload = rng.integers(0, 256, n_imtr * 866, dtype=np.uint8)
    payload[payload == 0xEB] = 0                      # no image-frame signatures: stage 1g finds nothing to write
    for k_restart in (700, 2200):
        imtr = synth.imtr_frames(payload, chid=0x22)
        imtr[k_restart, 4:8] = 0
        synth.refresh_imtr_crc(imtr[k_restart:k_restart + 1])
        buf = synth.aos_frames(imtr.reshape(-1)).reshape(-1)
        off, _ = oracle.aos_scan(buf)
        imdt_o, st = oracle.imtr_deframe(buf, off)
        work = tmp_path / f"out{k_restart}"
        work.mkdir()
        src = tmp_path / "KEL_MN200_20220316_120309_1.DAT"
        buf.tofile(str(src))
        assert REF.ref_auxsep(str(src).encode(), str(work).encode()) == 0
        name = [n for n in os.listdir(work) if n.endswith(".IMDT")][0]
        ref = np.fromfile(str(work / name), np.uint8)
        pre, post = (k_restart + 1) * 866, (n_imtr - k_restart - 1) * 866
        assert st[8] == 2 and imdt_o.size == post and np.array_equal(imdt_o, payload[pre:])
        assert np.array_equal(ref[:post], imdt_o)                       # the reference's file starts with the oracle's stream
        if post >= pre:
            assert ref.size == post                                     # ... and is nothing else when the new data is the longer part
        else:
            tail = ref[post:]
            nz = np.flatnonzero(tail)
            assert ref.size == pre and nz.size and nz[0] > tail.size - 8192
            assert np.array_equal(tail[nz[0]:], payload[post + nz[0]:pre])   # the old stream's last buffer, flushed after the truncation


def test_oracle_stage1_live_against_the_reference_on_damaged_downlinks(tmp_path):
    """three random downlinks at the reference geometry (3 image frames each) with damaged image-transfer frames (CRC,
    signature, tail, type), false sync words inside payloads, empty / bad-CRC / bad-inject AOS frames, a junk prefix and
    byte slips go through AuxSeparator::Separate (compiled unmodified) and through the oracle: same IMDT, AUX, PAN.RAW and
    MSS.RAW bytes.  (50 such cases, with random damage of every kind at once, were run when this test was written: all identical;
    mid-stream restarts are covered by the test above.)"""
    so = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_oip.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_oip.so not built (no /root/reference here)")
    import ctypes as C
    REF = C.CDLL(so)
    REF.ref_auxsep.argtypes = [C.c_char_p, C.c_char_p]
    rng = np.random.default_rng(2026)
    tc, tl = 1536, 256
    sync = np.frombuffer(synth.AOS_SYNC, np.uint8)
    frames_seen = 0
    for it in range(3):
        imdt, _ = synth.make_imdt(3, tc, tl, seed=int(rng.integers(1 << 30)), skip_seqs={2} if it == 1 else set())
        imtr = synth.imtr_frames(imdt, chid=int(rng.choice([0x11, 0x22])))
        k = int(rng.integers(0, imtr.shape[0] // 3))            # one damaged frame inside the first image frame
        kind = it % 3
        if kind == 0:
            imtr[k, int(rng.integers(10, 876))] ^= 0x08
        elif kind == 1:
            imtr[k, 880] ^= 0x01
        else:
            imtr[k, 9] = 0x33
            synth.refresh_imtr_crc(imtr[k:k + 1])
        aos = synth.aos_frames(imtr.reshape(-1)).copy()
        na = aos.shape[0]
        for _ in range(3):                                       # false sync words inside payloads
            q, p = int(rng.integers(0, na)), int(rng.integers(20, 880))
            aos[q, p:p + 4] = sync
            crc = synth.crc16_rows(aos[q:q + 1, 4:894])
            aos[q, 894], aos[q, 895] = crc[0] >> 8, crc[0] & 0xFF
        buf = synth.build_aos_file(aos, empty_every=int(rng.integers(50, 2000)), bad_crc_at=set(int(x) for x in rng.integers(0, na, 3)),
                                   bad_inject_at=set(int(x) for x in rng.integers(0, na, 2)), prefix=bytes(int(rng.integers(0, 40)))).copy()
        if it == 2:                                              # a byte slip late in the file: the cadence of the last frame is lost
            p = int(buf.size * 0.9)
            buf = np.concatenate([buf[:p], rng.integers(0, 256, 3, dtype=np.uint8), buf[p:]])
        buf = np.ascontiguousarray(buf)
        off, cnt = oracle.aos_scan(buf)
        imdt_o, st = oracle.imtr_deframe(buf, off)
        n, aux, pan, mss, fst = oracle.image_frames(imdt_o, tc, tl)
        work = tmp_path / f"out{it}"
        work.mkdir()
        src = tmp_path / "KEL_MN200_20220316_120309_1.DAT"
        buf.tofile(str(src))
        assert REF.ref_auxsep(str(src).encode(), str(work).encode()) == 0
        got = {}
        for nm in os.listdir(work):
            for ext in (".IMDT", ".AUX", ".PAN.RAW", ".MSS.RAW"):
                if nm.endswith(ext):
                    got[ext] = np.fromfile(str(work / nm), np.uint8)
        assert np.array_equal(got[".IMDT"], imdt_o), it
        for ext, arr in ((".AUX", aux), (".PAN.RAW", pan), (".MSS.RAW", mss)):
            assert np.array_equal(got[ext], np.ascontiguousarray(arr).view(np.uint8).reshape(-1)), (it, ext)
        frames_seen += n
    assert frames_seen >= 3
