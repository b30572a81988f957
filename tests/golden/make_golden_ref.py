"""Golden fixtures produced by RUNNING THE REFERENCE'S OWN CODE in this container
(oracle/_ref/libref_oip.so = reference headers compiled unmodified against stand-in third-party headers,
see oracle/ref_oip_shim.cpp).  Needs /root/reference; the fixtures travel, the .so does not have to.

  ref_auxsep.npz     AuxSeparator::Separate() on a synthetic AOS downlink (reference geometry, 3 image
                     frames, empty / bad-CRC / bad-inject AOS frames, a false sync word, a corrupted IMTR
                     frame that makes one image frame incomplete, a sequence gap): sha256 of the IMDT,
                     AUX, PAN.RAW and MSS.RAW files the reference wrote
  ref_prestitch.npz  Stitcher::PreStitch() on a 32768-line x 12288-px strip (two 30000-row sections, stale
                     bottom rows), dY < 0 and dY > 0: per-1024-row digests + boundary rows

  ref_bandalign.npz  PreProcessor::LoadMSS + DoRRC4MSS + DoInterBandAlignment (ref preproc.h:56-80, :202-222, :351-468)
                     at the reference geometry (4 x 3072-px bands): 7000 lines in 3000/520 sections with and without
                     -k / line offset / RRC, and 21600 lines in the default 20000/520 sections: per-512-row digests
  ref_rrc_csv.npz    IMO::LoadRRCParamFile (ref imageop.h:140-192) on well-formed and malformed CSV texts: the parsed
                     doubles or the fact that the reference threw

Run:  python tests/golden/make_golden_ref.py [auxsep] [prestitch] [bandalign] [rrccsv]
"""
import ctypes as C
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from opticalimageprocessor_b200 import synth  # noqa: E402

oracle.build()
REF = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_oip.so"))
REF.ref_auxsep.argtypes = [C.c_char_p, C.c_char_p]
REF.ref_prestitch.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_char_p]
W = REF.ref_pixels_per_line()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def auxsep_input(seed=20261018):
    """the synthetic downlink both the reference and the oracle/GPU tests parse"""
    tc, tl = 1536, 256
    imdt, _ = synth.make_imdt(4, tc, tl, seed=seed, skip_seqs={3})         # frames 1,2,4 -> gap at 3
    imtr = synth.imtr_frames(imdt, chid=0x22)
    frame_bytes = 192 * tl + 40 * tc * tl * 2 + 172
    k = (frame_bytes + frame_bytes // 2) // 866                              # an IMTR frame inside image frame 2
    imtr[k, 400] ^= 0x40                                                     # bad CRC -> dropped -> frame 2 incomplete
    aos = synth.aos_frames(imtr.reshape(-1)).copy()
    aos[77, 300:304] = np.frombuffer(synth.AOS_SYNC, np.uint8)              # false sync inside a payload
    crc = synth.crc16_rows(aos[77:78, 4:894])
    aos[77, 894], aos[77, 895] = crc[0] >> 8, crc[0] & 0xFF
    return synth.build_aos_file(aos, empty_every=997, bad_crc_at={5, 40000}, bad_inject_at={123}, prefix=b"\x00" * 13)


def make_auxsep():
    buf = auxsep_input()
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "KEL_MN200_20220316_120309_1.DAT")
        buf.tofile(src)
        work = os.path.join(d, "out")
        os.mkdir(work)
        rc = REF.ref_auxsep(src.encode(), work.encode())
        assert rc == 0, rc
        names = sorted(os.listdir(work))
        print("reference wrote:", names)
        stem = [n for n in names if n.endswith(".IMDT")][0][:-5]
        out = {"imdt_name": stem + ".IMDT"}
        for ext, key in [(".IMDT", "imdt"), (".AUX", "aux"), (".PAN.RAW", "pan"), (".MSS.RAW", "mss")]:
            a = np.fromfile(os.path.join(work, stem + ext), np.uint8)
            out[key + "_sha256"] = sha(a)
            out[key + "_bytes"] = a.size
    np.savez(os.path.join(HERE, "ref_auxsep.npz"), input_sha256=sha(buf), input_bytes=buf.size, **out)
    print(out)


def prestitch_input(rows=32768, seed=77):
    return np.random.default_rng(seed).integers(0, 65536, (rows, W), dtype=np.uint16)


def make_prestitch():
    rows = 32768
    src = prestitch_input(rows)
    res = {}
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "SYN_PAN-2.RRC.RAW")
        src.tofile(p)
        for tag, (dx, dy) in {"neg": (1.37, -2.61), "pos": (-0.83, 3.19)}.items():
            work = os.path.join(d, "w" + tag)
            os.mkdir(work)
            rc = REF.ref_prestitch(p.encode(), p.encode(), dx, dy, work.encode())
            assert rc == 0, rc
            out = np.fromfile(os.path.join(work, "SYN_PAN-2.RRC.PRESTT.RAW"), np.uint16).reshape(-1, W)
            assert out.shape[0] == rows
            res[tag + "_shift"] = np.array([dx, dy])
            res[tag + "_block_sha"] = np.array([sha(out[i:i + 1024]) for i in range(0, rows, 1024)])
            keep = list(range(0, 8)) + list(range(29990, 30010)) + list(range(rows - 12, rows))
            res[tag + "_rows_idx"] = np.array(keep)
            res[tag + "_rows"] = out[keep]
            print(tag, "done")
    np.savez_compressed(os.path.join(HERE, "ref_prestitch.npz"), rows=rows, seed=77, **res)


BA_CX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
BA_CY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8 * (b - 1.5)] for b in range(4)]
BA_CASES = {  # tag: (lines, lines_per_section, line_offset, overlap, keep_leading, do_rrc, seed)
    "s3000": (7000, 3000, 0, 520, False, True, 31),
    "s3000_keep_off_norrc": (7000, 3000, 100, 520, True, False, 32),
    "default": (21600, 20000, 0, 520, False, True, 33),
}


def bandalign_input(lines, seed):
    """MSS strip: `lines` lines of 4 x 3072 px side by side (ref preproc.h:62-75), full-range 16-bit noise"""
    return np.random.default_rng(seed).integers(0, 65536, (lines, 12288), dtype=np.uint16)


def make_bandalign():
    REF.ref_band_align.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    res = {}
    for tag, (lines, lps, off, ov, keep, do_rrc, seed) in BA_CASES.items():
        mss = bandalign_input(lines, seed)
        with tempfile.TemporaryDirectory() as d:
            mp = os.path.join(d, "SYN_CMOS-1.MSS.RAW")
            pp = os.path.join(d, "SYN_CMOS-1.PAN.RAW")
            mss.tofile(mp)
            with open(pp, "wb") as f:                       # only its size is looked at (ref preproc.h:563-566): sparse
                f.truncate(4 * mss.nbytes)
            rrc = []
            for b in range(4):
                q = os.path.join(d, f"rrc_b{b + 1}.csv")
                synth.write_rrc_csv(q, synth.rrc_coeffs(3072, 300 + b))
                rrc.append(q.encode())
            work = os.path.join(d, "out")
            os.mkdir(work)
            cx = (C.c_double * 8)(*[v for r in BA_CX for v in r])
            cy = (C.c_double * 12)(*[v for r in BA_CY for v in r])
            rc = REF.ref_band_align(pp.encode(), mp.encode(), (C.c_char_p * 4)(*rrc), int(do_rrc), cx, cy, lps, off, ov, int(keep), work.encode())
            assert rc == 0, rc
            out = np.fromfile(os.path.join(work, "SYN_CMOS-1.MSS.ALIGNED.TIFF"), np.uint16).reshape(-1, 3072, 4)
        rows = lines - off - (0 if keep else ov)
        assert out.shape[0] == rows, (out.shape, rows)
        res[tag + "_params"] = np.array([lines, lps, off, ov, int(keep), int(do_rrc), seed])
        res[tag + "_block_sha"] = np.array([sha(out[i:i + 512]) for i in range(0, rows, 512)])
        res[tag + "_rows"] = rows
        print(tag, out.shape, "nonzero rows:", int((out.reshape(rows, -1) != 0).any(axis=1).sum()))
    np.savez_compressed(os.path.join(HERE, "ref_bandalign.npz"), cX=np.array(BA_CX), cY=np.array(BA_CY), **res)


RRC_CSV_TEXTS = {
    "plain": "1\n4\n0\n1.000000000000 , 0.500000000000\n0.987654321012 , 7.250000000000\n1.05 , 0\n0.95,8\n",
    "spaces_crlf": "1\r\n3\r\n0\r\n   1.5   ,   2.5  \r\n\t0.25,\t-3e-3\r\n1e0 , 1E1\r\n",
    "no_final_newline": "1\n2\n0\n1.25 , 0.75\n0.5 , 0.125",
    "count_mismatch_header": "1\n5\n0\n1 , 0\n1 , 0\n1 , 0\n1 , 0\n",
    "too_few_rows": "1\n4\n0\n1 , 0\n1 , 0\n",
    "bad_row": "1\n2\n0\n1 , 0\nfoo , 1\n",
    "missing_comma": "1\n2\n0\n1 0\n1 , 0\n",
    "blank_trailing_line": "1\n2\n0\n1 , 0\n2 , 1\n\n",
    "header_only": "1\n",
    "first_lines_arbitrary": "hello\n2 trailing text\nworld\n3.5 , 4.5 extra\n-1e300 , 1e-300\n",
}
RRC_CSV_EXPECT = {"plain": 4, "spaces_crlf": 3, "no_final_newline": 2, "count_mismatch_header": 4, "too_few_rows": 4, "bad_row": 2,
                  "missing_comma": 2, "blank_trailing_line": 2, "header_only": 2, "first_lines_arbitrary": 2}


def make_rrccsv():
    REF.ref_load_rrc_csv.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, text in RRC_CSV_TEXTS.items():
            p = os.path.join(d, tag + ".csv")
            with open(p, "wb") as f:
                f.write(text.encode())
            n = RRC_CSV_EXPECT[tag]
            kb = (C.c_double * (2 * n))()
            rc = REF.ref_load_rrc_csv(p.encode(), n, kb)
            res[tag + "_ok"] = rc == 0
            res[tag + "_kb"] = np.array(list(kb)) if rc == 0 else np.zeros(0)
            print(tag, rc, list(kb)[:4] if rc == 0 else "")
        rc = REF.ref_load_rrc_csv(os.path.join(d, "does_not_exist.csv").encode(), 2, (C.c_double * 4)())
        res["missing_file_ok"] = rc == 0
    np.savez(os.path.join(HERE, "ref_rrc_csv.npz"), **res)


if __name__ == "__main__":
    what = sys.argv[1:] or ["auxsep", "prestitch", "bandalign", "rrccsv"]
    if "auxsep" in what:
        make_auxsep()
    if "prestitch" in what:
        make_prestitch()
    if "bandalign" in what:
        make_bandalign()
    if "rrccsv" in what:
        make_rrccsv()
