"""Live check of the oracle's PreStitch against THE REFERENCE ITSELF (oracle/_ref/libref_oip.so = stitcher.h / imageop.h compiled
unmodified) on shifts the golden file does not hold: zero, integer, sub-pixel, large and tiny ones, at the reference geometry
(12288-px lines, 30000-row sections, > 32767 lines).  Needs /root/reference (build container only); ~45 s per shift, which is why
it is a tool and not part of the CPU suite.  Result at the end of round 2 (32768 lines, all nine shifts): IDENTICAL.

    python tools/ref_live_prestitch.py [lines]
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
REF = C.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libref_oip.so'))
REF.ref_prestitch.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_char_p]
W = 12288
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
src = np.random.default_rng(5).integers(0, 65536, (rows, W), dtype=np.uint16)
cases = [(0.0, 0.0), (2.0, -3.0), (-1.0, 4.0), (0.5, 0.25), (-0.25, -0.5), (7.9, 31.4), (-3.3, -40.7), (1.37, 1e-9), (0.0, -1e-9)]
with tempfile.TemporaryDirectory(dir='/dev/shm' if os.path.isdir('/dev/shm') else None) as d:
    p = os.path.join(d, "SYN_PAN-2.RRC.RAW"); src.tofile(p)
    for k, (dx, dy) in enumerate(cases):
        t0 = time.time()
        work = os.path.join(d, f"w{k}"); os.mkdir(work)
        rc = REF.ref_prestitch(p.encode(), p.encode(), dx, dy, work.encode())
        if rc != 0:
            print(dx, dy, "ref rc", rc); continue
        out = np.fromfile(os.path.join(work, "SYN_PAN-2.RRC.PRESTT.RAW"), np.uint16).reshape(-1, W)
        t1 = time.time()
        mine = oracle.prestitch_shift(src, dx, dy)
        neq = np.argwhere(out != mine)
        print((dx, dy), "ref", round(t1 - t0, 1), "s, oracle", round(time.time() - t1, 1), "s:", "IDENTICAL" if neq.size == 0 else f"{len(neq)} px differ, rows {sorted(set(neq[:,0].tolist()))[:12]}", flush=True)
        for f in os.listdir(work): os.remove(os.path.join(work, f))
