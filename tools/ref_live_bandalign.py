"""Live check of the oracle's band alignment against THE REFERENCE ITSELF (oracle/_ref/libref_oip.so = preproc.h compiled
unmodified: PreProcessor::LoadMSS + DoRRC4MSS + DoInterBandAlignment) on random polynomials and section geometries the golden
file does not hold: flat to steep coefficients of both signs, section lengths, overlaps, line offsets, keep-leading, RRC on / off.
Needs /root/reference (build container only).

    python tools/ref_live_bandalign.py [seed] [cases]
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from opticalimageprocessor_b200 import synth  # noqa: E402

REF = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_oip.so"))
REF.ref_band_align.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 6
bad = 0
for it in range(n_cases):
    t0 = time.time()
    lps = int(rng.integers(1600, 4000))
    ov = int(rng.integers(0, 600))
    off = int(rng.choice([0, 0, 7, 100]))
    keep = bool(rng.random() < 0.4)
    do_rrc = bool(rng.random() < 0.6)
    lines = int(rng.integers(lps + ov + off + 1600, 9000))
    scale = float(rng.choice([0.0, 0.3, 1.0, 3.0, 10.0]))
    cX = [[float(rng.uniform(-2, 2)), float(rng.uniform(-3e-4, 3e-4)) * scale] for _ in range(4)]
    cY = [[float(rng.uniform(-5, 5)), float(rng.uniform(-4e-4, 4e-4)) * scale, float(rng.uniform(-2e-8, 2e-8)) * scale] for _ in range(4)]
    mss = rng.integers(0, 65536, (lines, 12288), dtype=np.uint16)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        mp, pp = os.path.join(d, "SYN_CMOS-1.MSS.RAW"), os.path.join(d, "SYN_CMOS-1.PAN.RAW")
        mss.tofile(mp)
        with open(pp, "wb") as f:
            f.truncate(4 * mss.nbytes)
        rrc = []
        for b in range(4):
            q = os.path.join(d, f"rrc_b{b + 1}.csv")
            synth.write_rrc_csv(q, synth.rrc_coeffs(3072, 300 + b))
            rrc.append(q.encode())
        work = os.path.join(d, "out")
        os.mkdir(work)
        cx = (C.c_double * 8)(*[v for r in cX for v in r])
        cy = (C.c_double * 12)(*[v for r in cY for v in r])
        rc = REF.ref_band_align(pp.encode(), mp.encode(), (C.c_char_p * 4)(*rrc), int(do_rrc), cx, cy, lps, off, ov, int(keep), work.encode())
        ref = np.fromfile(os.path.join(work, "SYN_CMOS-1.MSS.ALIGNED.TIFF"), np.uint16).reshape(-1, 3072, 4) if rc == 0 else None
    planes = oracle.mss_split(mss)
    if do_rrc:
        planes = [oracle.rrc(p, synth.rrc_coeffs(3072, 300 + b)) for b, p in enumerate(planes)]
    n, out = oracle.band_align(planes, np.array(cX), np.array(cY), lines_per_section=lps, line_offset=off, overlap=ov, keep_leading=keep)
    what = dict(lines=lines, lps=lps, ov=ov, off=off, keep=keep, rrc=do_rrc, scale=scale)
    if ref is None:
        print(it, what, "reference rc", rc)
        bad += 1
        continue
    same = ref.shape == out.shape and np.array_equal(ref, out)
    if not same:
        bad += 1
        neq = np.argwhere(ref != out) if ref.shape == out.shape else None
        print(it, what, "DIFFER", ref.shape, out.shape, None if neq is None else (len(neq), sorted(set(neq[:, 0].tolist()))[:10]))
    else:
        print(it, what, "IDENTICAL", round(time.time() - t0, 1), "s", flush=True)
print("done", bad, "bad")
