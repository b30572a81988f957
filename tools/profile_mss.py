"""ncu driver: oip_band_align_merge on a reference-geometry MSS strip (4 x 3072 px, 16384 lines, LE16 + RRC)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops, synth
ctx = ops.Context(0)
W, wb, lines = 12288, 3072, int(os.environ.get("LINES", 16384))
mss = torch.randint(0, 4096, (lines, W), device="cuda", dtype=torch.int32).to(torch.uint16)
kbs = [torch.from_numpy(synth.rrc_coeffs(wb, 20 + i)).cuda() for i in range(4)]
cX = [[0.8 + 0.1 * i, -1.5e-4 * (i + 1)] for i in range(4)]
cY = [[-3.2 + i, 2e-4 * (i + 1), -1e-8 * (i - 1.5)] for i in range(4)]
out = torch.zeros((lines - 520, wb, 4), dtype=torch.uint16, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    if i == 3: e0.record()
    ops.band_align(ctx, mss, wb, kbs, cX, cY, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"band_align {lines} x {W}: {ms:.3f} ms  {lines*W/ms/1e6:.1f} Gpx/s  {(lines*W*2+out.numel()*2)/ms/1e6:.0f} GB/s")
