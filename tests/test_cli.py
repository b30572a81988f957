"""The drop-in boundary at process level: OpticalImageProcessor CLI (C++ host over the C ABI) keeps the
reference's argv grammar, exit codes (ref main.cpp:260-267, :333-342) and cwd-relative output names
(ref imageop.h:99-108).  Exit-code tests need no GPU; the file-flow tests compare the files the CLI
writes with the files THE REFERENCE ITSELF wrote (tests/golden/ref_*.npz) and with the oracle."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from opticalimageprocessor_b200 import build, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cli():
    build.build()
    assert os.path.exists(build.CLI_BIN)
    return build.CLI_BIN


def run(cli, args, cwd):
    env = dict(os.environ, LOGFILE=os.path.join(cwd, "test.log"))
    return subprocess.run([cli] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)


def sha_file(p):
    h = hashlib.sha256()
    with open(p, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def test_exit_codes_and_grammar(cli, tmp_path):
    d = str(tmp_path)
    assert run(cli, ["-v"], d).returncode == 255 and run(cli, ["--version"], d).stdout.strip() == "1.1"  # app.exit(e)+255
    assert run(cli, ["--help"], d).returncode == 255
    assert run(cli, ["stitch", "--image1", "a.RAW"], d).returncode == 106           # CLI::RequiredError
    assert run(cli, ["auxsep", "/nonexistent"], d).returncode == 105                # CLI::ExistingFile validator
    assert run(cli, ["--bogus"], d).returncode == 109                               # CLI::ExtrasError
    assert run(cli, ["stitch", "--image1", "a.RAW", "--image2", "b.RAW", "-c", "1"], d).returncode == 105  # fold < 2
    r = run(cli, [], d)                                                              # usage_error -> 254, ref main.cpp:333-335
    assert r.returncode == 254 and "USAGE ERROR: RRC parameter file of all MSS Bands needed." in r.stdout
    open(os.path.join(d, "x.RAW"), "wb").write(b"\0" * 10)
    open(os.path.join(d, "y.TIFF"), "wb").write(b"\0" * 10)
    r = run(cli, ["stitch", "--image1", "x.RAW", "--image2", "y.TIFF", "-c", "200", "-o", "o.RAW"], d)
    assert r.returncode == 2 and "two images should be same type" in r.stdout      # std::exception -> 2, ref :336-338
    assert "two images should be same type" in open(os.path.join(d, "test.log")).read()  # LOGFILE, ref :324-328
    r = run(cli, ["auxsep", "x.RAW"], d)
    assert r.returncode == 2 and "unrecognized AOS file name pattern" in r.stdout   # ref aux_separator.h:208-213


@pytest.mark.gpu
def test_auxsep_files_equal_reference_run(cli, tmp_path):
    """same downlink the reference's AuxSeparator parsed: identical file names and bytes"""
    from test_reference_golden import _inputs
    g = np.load(os.path.join(GOLD, "ref_auxsep.npz"))
    d = str(tmp_path)
    src = os.path.join(d, "in")
    os.mkdir(src)
    p = os.path.join(src, "KEL_MN200_20220316_120309_1.DAT")
    _inputs()["auxsep_input"]().tofile(p)
    r = run(cli, ["auxsep", p], d)
    assert r.returncode == 0, r.stdout + r.stderr
    stem = str(g["imdt_name"])[:-5]
    for ext, key in [(".IMDT", "imdt"), (".AUX", "aux"), (".PAN.RAW", "pan"), (".MSS.RAW", "mss")]:
        f = os.path.join(d, stem + ext)
        assert os.path.exists(f), f"{stem + ext} missing; cwd has {os.listdir(d)}"
        assert os.path.getsize(f) == int(g[key + "_bytes"]) and sha_file(f) == str(g[key + "_sha256"]), ext
    # .IMDT shortcut (ref aux_separator.h:204-206,230): same products again from the intermediate file
    d2 = os.path.join(d, "again")
    os.mkdir(d2)
    r = run(cli, ["auxsep", os.path.join(d, stem + ".IMDT")], d2)
    assert r.returncode == 0
    assert sha_file(os.path.join(d2, stem + ".PAN.RAW")) == str(g["pan_sha256"])


@pytest.mark.gpu
def test_prestitch_and_stitch_task_flow(cli, tmp_path, oracle_mod):
    """DOC/Usage.txt steps 1-2 on a 32768-line strip: file names, RRC files, PRESTT file, stitched RAW"""
    from test_reference_golden import _inputs, _check_prestitch
    g = np.load(os.path.join(GOLD, "ref_prestitch.npz"))
    d = str(tmp_path)
    W, rows = 12288, int(g["rows"])
    src = _inputs()["prestitch_input"](rows, int(g["seed"]))
    p2 = os.path.join(d, "SYN_PAN-2.RAW")
    src.tofile(p2)
    dx, dy = [float(v) for v in g["neg_shift"]]
    # --no-rrc: PRESTT of the raw file must equal what the reference's PreStitch wrote for this input
    sec = ["-s", "2", "-l", "16000"]           # defaults (10 x 16000 lines) need a 160000-line strip, ref stitcher.h:60-77
    r = run(cli, ["prestitch", "--pan1", p2, "--pan2", p2, "--no-rrc", f"--dx={dx}", f"--dy={dy}"] + sec, d)
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(os.path.join(d, "SYN_PAN-2.PRESTT.RAW"), np.uint16).reshape(rows, W)
    _check_prestitch(out, g, "neg")
    # without --dx/--dy the offsets are estimated (Stitcher::CalcSttParameters, ref stitcher.h:148-201); the overlap
    # columns of this pair are unrelated noise -> no valid section -> std::runtime_error -> exit code 2 (ref :190-192)
    r = run(cli, ["prestitch", "--pan1", p2, "--pan2", p2, "-c"] + sec, d)
    assert r.returncode == 2 and "No valid delta value found" in (r.stdout + r.stderr)
    # a pair that does overlap: -c prints the table and the means, which must match the reference loop on cv2.phaseCorrelate
    import cv2
    from test_phasecorr_cpu import _pair
    sa, sb = _pair(rows, 200, 1.37, -2.61, seed=11)
    o1, o2 = src.copy(), src[::-1].copy()
    o1[:, W - 200:] = sa
    o2[:, :200] = sb
    q1, q2 = os.path.join(d, "OVL_PAN-1.RAW"), os.path.join(d, "OVL_PAN-2.RAW")
    o1.tofile(q1); o2.tofile(q2)
    r = run(cli, ["prestitch", "--pan1", q1, "--pan2", q2, "-c", "-e", "2"] + sec, d)
    assert r.returncode == 0, r.stdout + r.stderr
    import re
    m = re.search(r"dx: (-?[0-9.]+), dy: (-?[0-9.]+), r: (-?[0-9.]+)", r.stdout)
    assert m and "Total 2 valid delta value pairs found" in r.stdout

    def cvcorr(s1, s2):
        (x, y), rr = cv2.phaseCorrelate(s1, s2)
        return x, y, rr
    _, mean = oracle_mod.stt_parameters(o1, o2, overlap_cols=200, edge_cols=2, sections=2, lines_per_section=16000, correlate=cvcorr)
    assert all(abs(float(m.group(k + 1)) - mean[k]) <= 2e-3 for k in range(3))
    assert run(cli, ["prestitch", "--pan1", p2, "--pan2", p2, "--dx=1", "--dy=1"], d).returncode == 2   # too few lines for 10 x 16000
    # with RRC (default): <stem>.RRC.RAW for both and <stem>.RRC.PRESTT.RAW, first/last rows vs oracle
    kb1, kb2 = synth.rrc_coeffs(W, 1), synth.rrc_coeffs(W, 2)
    synth.write_rrc_csv(os.path.join(d, "PAN-1.csv"), kb1)
    synth.write_rrc_csv(os.path.join(d, "PAN-2.csv"), kb2)
    p1 = os.path.join(d, "SYN_PAN-1.RAW")
    src[::-1].tofile(p1)
    r = run(cli, ["prestitch", f"--pan1={p1}", f"--pan2={p2}", "--rrc1", "PAN-1.csv", "--rrc2", "PAN-2.csv", "--dx", str(dx), "--dy", str(dy)] + sec, d)
    assert r.returncode == 0, r.stdout + r.stderr
    sl = slice(0, 64)
    rrc1 = np.fromfile(os.path.join(d, "SYN_PAN-1.RRC.RAW"), np.uint16).reshape(rows, W)
    assert np.array_equal(rrc1[sl], oracle_mod.rrc(src[::-1][sl], kb1))
    rrc2 = np.fromfile(os.path.join(d, "SYN_PAN-2.RRC.RAW"), np.uint16).reshape(rows, W)
    assert np.array_equal(rrc2[-64:], oracle_mod.rrc(src[-64:], kb2))
    pre = np.fromfile(os.path.join(d, "SYN_PAN-2.RRC.PRESTT.RAW"), np.uint16).reshape(rows, W)
    want_top = oracle_mod.prestitch_shift(np.ascontiguousarray(rrc2[:600]), dx, dy)   # interior rows do not depend on the strip length
    assert np.array_equal(pre[8:500], want_top[8:500])
    # step 2: stitch the two RAW products (ref imageop.h:340-355)
    r = run(cli, ["stitch", "--image1=SYN_PAN-1.RRC.RAW", "--image2=SYN_PAN-2.RRC.PRESTT.RAW", "--fold-cols=200", "-o", "stitched-PAN.RAW"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    st = np.fromfile(os.path.join(d, "stitched-PAN.RAW"), np.uint16).reshape(rows, 2 * (W - 100))
    assert np.array_equal(st[:, :W - 100], rrc1[:, :W - 100]) and np.array_equal(st[:, W - 100:], pre[:, 100:])
    # no -o: single-band TIFF under the reference's default name (ref imageop.h:299-303, GTiff :316-328)
    r = run(cli, ["stitch", "--image1=SYN_PAN-1.RRC.RAW", "--image2=SYN_PAN-2.RRC.PRESTT.RAW", "--fold-cols=200"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    import cv2
    os.environ.setdefault("OPENCV_IO_MAX_IMAGE_PIXELS", str(1 << 40))
    tif = cv2.imread(os.path.join(d, f"stitched_{2 * (W - 100)}n16b.TIFF"), cv2.IMREAD_UNCHANGED)
    assert tif is not None and np.array_equal(tif, st)


@pytest.mark.gpu
def test_default_action_mss(cli, tmp_path, oracle_mod):
    d = str(tmp_path)
    W, wb, lines = 12288, 3072, 2100
    rng = np.random.default_rng(3)
    mss = rng.integers(0, 4096, (lines, W), dtype=np.uint16)
    mss.tofile(os.path.join(d, "SYN_MSS-1.RAW"))
    np.zeros((lines * 4, W), np.uint16).tofile(os.path.join(d, "SYN_PAN-1.RRC.RAW"))
    kbs = [synth.rrc_coeffs(wb, 10 + b) for b in range(4)]
    args = ["--pan=SYN_PAN-1.RRC.RAW", "--mss=SYN_MSS-1.RAW"]
    for b in range(4):
        synth.write_rrc_csv(os.path.join(d, f"MSS-1.B{b + 1}.csv"), kbs[b])
        args += [f"--rrc-msb{b + 1}", f"MSS-1.B{b + 1}.csv"]
    cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
    cY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8 * (b - 1.5)] for b in range(4)]
    with open(os.path.join(d, "poly.txt"), "w") as f:
        for b in range(4):
            f.write("%r %r %r %r %r\n" % (cX[b][0], cX[b][1], cY[b][0], cY[b][1], cY[b][2]))
    # without --poly the coefficients are estimated (CalcInterBandCorrelation); this PAN has 8400 lines, too few for the
    # default 5 sections x 16000 lines -> std::invalid_argument -> exit code 2 (ref preproc.h:234-237, main.cpp:336-339)
    r = run(cli, args, d)
    assert r.returncode == 2 and "too many sections" in (r.stdout + r.stderr)
    r = run(cli, args + ["--poly", "poly.txt", "--aligned-raw"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    planes = [oracle_mod.rrc(p, k) for p, k in zip(oracle_mod.mss_split(mss), kbs)]
    n, want = oracle_mod.band_align(planes, cX, cY)
    got = np.fromfile(os.path.join(d, "SYN_MSS-1.ALIGNED.RAW"), np.uint16).reshape(lines - 520, wb, 4)
    assert n == lines - 520 and np.array_equal(got[:n], want[:n])
    # the reference's product is <stem>.ALIGNED.TIFF written by cv::imwrite (ref preproc.h:167-185): read it the way the
    # reference reads such files (cv::imread IMREAD_UNCHANGED, ref imageop.h:392) -> the CV_16UC4 memory image
    import cv2
    tif = cv2.imread(os.path.join(d, "SYN_MSS-1.ALIGNED.TIFF"), cv2.IMREAD_UNCHANGED)
    assert tif is not None and tif.shape == (lines - 520, wb, 4) and np.array_equal(tif[:n], want[:n])
    # stitch two 4-channel TIFFs (IMO::StitchTiff, ref imageop.h:365-457; GDAL variant with band map :460-567)
    os.rename(os.path.join(d, "SYN_MSS-1.ALIGNED.TIFF"), os.path.join(d, "L.TIFF"))
    cv2.imwrite(os.path.join(d, "R.TIFF"), tif[:, ::-1].copy())   # a libtiff-written input with cv::imwrite's defaults (LZW, predictor 2)
    r = run(cli, ["stitch", "--image1=L.TIFF", "--image2=R.TIFF", "--fold-cols=50"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    st = cv2.imread(os.path.join(d, "stitched.TIFF"), cv2.IMREAD_UNCHANGED)                          # default name, ref :373-375
    ref_st = oracle_mod.stitch_concat_c4([tif, tif[:, ::-1].copy()], 25)
    assert np.array_equal(st, ref_st)
    r = run(cli, ["stitch", "--image1=L.TIFF", "--image2=R.TIFF", "--fold-cols=50", "-g", "-m", "3,2,1,4", "-o", "g.TIFF"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    g = cv2.imread(os.path.join(d, "g.TIFF"), cv2.IMREAD_UNCHANGED)
    # GDAL band b = memory channel map[b]-1 (ref :529); cv2.imread swaps samples 0/2 of what is in the file
    file_order = oracle_mod.stitch_concat_c4([tif, tif[:, ::-1].copy()], 25, [3, 2, 1, 4])
    assert np.array_equal(g, file_order[:, :, [2, 1, 0, 3]])
    assert run(cli, ["stitch", "--image1=L.TIFF", "--image2=R.TIFF", "--fold-cols=50", "-o", "x.RAW"], d).returncode == 2   # ref :376-380


def test_downlink_extension_grammar(cli, tmp_path):
    """`downlink` (extension: auxsep -> prestitch -> stitch for the PAN product in one fused pass) follows the conventions of
    the reference's sub-commands: required options, file validators, fold-cols check, RRC files unless --no-rrc"""
    d = str(tmp_path)
    open(os.path.join(d, "a.DAT"), "wb").write(b"\0" * 4096)
    assert run(cli, ["downlink", "--aos1", "a.DAT"], d).returncode == 106                                   # CLI::RequiredError
    assert run(cli, ["downlink", "--aos1", "a.DAT", "--aos2", "nope.DAT", "--dx", "1", "--dy", "1", "-c", "200"], d).returncode == 105
    assert run(cli, ["downlink", "--aos1", "a.DAT", "--aos2", "a.DAT", "--dx", "1", "--dy", "1", "-c", "1"], d).returncode == 105
    assert run(cli, ["downlink", "--aos1", "a.DAT", "--aos2", "a.DAT", "--dx", "x", "--dy", "1", "-c", "200", "--no-rrc"], d).returncode == 104  # CLI::ConversionError
    r = run(cli, ["downlink", "--aos1", "a.DAT", "--aos2", "a.DAT", "--dx", "1", "--dy", "1", "-c", "200"], d)
    assert r.returncode == 2 and "open RRC Param file failed" in r.stdout


@pytest.mark.gpu
def test_downlink_extension_equals_the_three_step_chain(cli, tmp_path, oracle_mod):
    """both CMOS downlinks in, stitched raster out: byte-identical to the oracle's chain aos_scan -> imtr_deframe ->
    image_frames -> RRC + shift + stitch on the downlink the reference's own auxsep golden was made from (reference geometry,
    a complete frame, two incomplete ones and zero-filled gap frames); the same file serves as CMOS-1 and CMOS-2
    (measured once by hand: profiles/r02k_cli_downlink.log)"""
    from test_reference_golden import _inputs
    W = 12288
    d = str(tmp_path)
    buf = _inputs()["auxsep_input"]()
    p = os.path.join(d, "KEL_MN200_20220316_120309_1.DAT")
    buf.tofile(p)
    kb1, kb2 = synth.rrc_coeffs(W, 1), synth.rrc_coeffs(W, 2)
    synth.write_rrc_csv(os.path.join(d, "PAN-1.csv"), kb1)
    synth.write_rrc_csv(os.path.join(d, "PAN-2.csv"), kb2)
    off, _ = oracle_mod.aos_scan(buf)
    imdt, _ = oracle_mod.imtr_deframe(buf, off)
    n, _, pan, _, _ = oracle_mod.image_frames(imdt, 1536, 256)
    dx, dy, fold = 1.37, -2.61, 200
    want = oracle_mod.pan_pipeline([pan, pan], [kb1, kb2], [0, dx], [0, dy], fold // 2, 30000, 32767)
    r = run(cli, ["downlink", "--aos1", p, "--aos2", p, "--rrc1", "PAN-1.csv", "--rrc2", "PAN-2.csv", "--dx", str(dx), "--dy", str(dy),
                  "-c", str(fold), "-o", "out.RAW"], d)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(os.path.join(d, "out.RAW"), np.uint16).reshape(-1, want.shape[1])
    assert got.shape == want.shape == (n * 1024, 2 * W - fold) and np.array_equal(got, want)
    assert not [f for f in os.listdir(d) if f.upper().endswith((".IMDT", ".AUX", ".PAN.RAW", ".MSS.RAW", ".RRC.RAW", ".PRESTT.RAW"))]  # no intermediate files
