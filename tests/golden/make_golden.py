"""Generates the committed golden fixtures from reference-backed sources, in THIS container:
  crc16.npz     : CRC values from the reference's own CRC.h (oracle/_ref/libref_crc.so)
  remap_cv2.npz : cv2.remap(INTER_CUBIC, BORDER_CONSTANT) outputs -- the library call the reference
                  makes at imageop.h:258 / preproc.h:453 (cv2 4.13.0 here)
  cubic_tab.npz : the 32x4 bicubic weight table, each entry confirmed against cv2 via impulse images
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

oracle.build()
ref = oracle.ref_crc_lib()
assert ref is not None, "needs /root/reference (oracle/_ref/libref_crc.so)"

rng = np.random.default_rng(20261018)
msgs = rng.integers(0, 256, (24, 890), dtype=np.uint8)
msgs[0] = 0
msgs[1] = 0xFF
msgs[2, :9] = np.frombuffer(b"123456789", np.uint8)
crcs = np.array([ref.ref_crc16_ccitt_false(np.ascontiguousarray(m), m.size) for m in msgs], np.uint16)
np.savez_compressed(os.path.join(HERE, "crc16.npz"), msgs=msgs, crcs=crcs, cv2_version=cv2.__version__)

H, W = 61, 200
src = rng.integers(0, 65536, (H, W), dtype=np.uint16)
shifts = np.array([(0, 0), (3, -2), (0.5, 0.5), (1.37, -2.61), (-1.984375, 4.015625), (0.015625, 0), (-5.3, 7.77),
                   (-0.83, 3.19)], np.float64)
outs = []
for dX, dY in shifts:
    mx = (np.arange(W)[None, :] + np.zeros((H, 1)) + dX).astype(np.float32)
    my = (np.arange(H)[:, None] + np.zeros((1, W)) + dY).astype(np.float32)
    outs.append(cv2.remap(src, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT))
np.savez_compressed(os.path.join(HERE, "remap_cv2.npz"), src=src, shifts=shifts, out=np.stack(outs),
                    cv2_version=cv2.__version__)

# weight table: extract w1d[f][k] from cv2 with a float impulse (out = wy[1]*wx[k'] etc.)
tab = oracle.cubic_tab()
imp = np.zeros((9, 9), np.float32)
imp[4, 4] = 1.0
for f in range(32):
    for k in range(4):
        # tap k of a pixel mapped to x = 4 - (k-1) + f/32 lands on the impulse with weight wx[k]; y weight = wy(0)[1] = 1
        mx = np.full((1, 1), 4 - (k - 1) + f / 32.0, np.float32)
        my = np.full((1, 1), 4.0, np.float32)
        got = cv2.remap(imp, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)[0, 0]
        assert got == np.float32(tab[0, 1]) * tab[f, k], (f, k, got, tab[f, k])
np.savez_compressed(os.path.join(HERE, "cubic_tab.npz"), tab=tab, cv2_version=cv2.__version__)
print("golden fixtures written to", HERE)
