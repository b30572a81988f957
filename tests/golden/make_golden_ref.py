"""Golden fixtures produced by RUNNING THE REFERENCE'S OWN CODE in this container
(oracle/_ref/libref_oip.so = reference headers compiled unmodified against stand-in third-party headers,
see oracle/ref_oip_shim.cpp).  Needs /root/reference; the fixtures travel, the .so does not have to.

  ref_auxsep.npz     AuxSeparator::Separate() on a synthetic AOS downlink (reference geometry, 3 image
                     frames, empty / bad-CRC / bad-inject AOS frames, a false sync word, a corrupted IMTR
                     frame that makes one image frame incomplete, a sequence gap): sha256 of the IMDT,
                     AUX, PAN.RAW and MSS.RAW files the reference wrote
  ref_prestitch.npz  Stitcher::PreStitch() on a 32768-line x 12288-px strip (two 30000-row sections, stale
                     bottom rows), dY < 0 and dY > 0: per-1024-row digests + boundary rows

Run:  python tests/golden/make_golden_ref.py [auxsep] [prestitch]
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


if __name__ == "__main__":
    what = sys.argv[1:] or ["auxsep", "prestitch"]
    if "auxsep" in what:
        make_auxsep()
    if "prestitch" in what:
        make_prestitch()
