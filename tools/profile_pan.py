"""minimal driver for ncu: the bench.py workload (3 CCD x 8192 px, BE16, fold 200), fused PAN kernels only, ROWS lines
(default: the C4 strip of bench.py; ROWS=32768 = C2).  3 warm-up + 2 timed launches."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from opticalimageprocessor_b200 import ops, synth, capi
ctx = ops.Context(0)
rows = int(os.environ.get("ROWS", bench.TOTAL_ROWS))
ccds = []
for i in range(bench.N_CCD):
    t = torch.empty((rows, bench.W), dtype=torch.uint16, device="cuda")
    capi.check(ctx.lib.oip_synth_strip_dn(ctx.h, t.data_ptr(), bench.W, rows, 0, bench.W, bench.SEED + i, 1))
    ccds.append(t)
kbs = [torch.from_numpy(synth.rrc_coeffs(bench.W, bench.SEED + 100 + i)).cuda() for i in range(bench.N_CCD)]
out = torch.empty((rows, ops.pan_out_width(bench.N_CCD, bench.W, bench.FOLD // 2)), dtype=torch.uint16, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = int(os.environ.get("ITERS", 2))
for i in range(3 + iters):
    if i == 3: e0.record()
    ops.pan_pipeline(ctx, ccds, kbs, bench.DX, bench.DY, bench.FOLD // 2, fmt=ops.FMT_BE16, out=out, check_error=False)
e1.record(); torch.cuda.synchronize()
capi.check(ctx.lib.oip_pan_check_error(ctx.h))
ms = e0.elapsed_time(e1) / iters
px = bench.N_CCD * bench.W * rows
if os.environ.get("CHECKSUM"):
    print("checksum", int(out.view(torch.int16).to(torch.int64).sum().item()), int(out.view(torch.int16)[::7, ::13].to(torch.int64).sum().item()))
print(f"fused PAN rows={rows}: {ms:.3f} ms  {px/ms/1e6:.1f} Gpx/s  {px*bench.algorithmic_bytes_per_px()/ms/1e6:.1f} GB/s")
