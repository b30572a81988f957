"""fused PAN on the bench geometry under different context options (stages per warp, rows per warp-tile); ROWS lines"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from opticalimageprocessor_b200 import ops, synth, capi
ctx = ops.Context(0)
rows = int(os.environ.get("ROWS", 131072))
ccds = []
for i in range(bench.N_CCD):
    t = torch.empty((rows, bench.W), dtype=torch.uint16, device="cuda")
    capi.check(ctx.lib.oip_synth_strip_dn(ctx.h, t.data_ptr(), bench.W, rows, 0, bench.W, bench.SEED + i, 1))
    ccds.append(t)
kbs = [torch.from_numpy(synth.rrc_coeffs(bench.W, bench.SEED + 100 + i)).cuda() for i in range(bench.N_CCD)]
out = torch.empty((rows, ops.pan_out_width(bench.N_CCD, bench.W, bench.FOLD // 2)), dtype=torch.uint16, device="cuda")
def run(tag):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(23):
        if i == 3: e0.record()
        ops.pan_pipeline(ctx, ccds, kbs, bench.DX, bench.DY, bench.FOLD // 2, fmt=ops.FMT_BE16, out=out, check_error=False)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
for st in (3, 4, 5, 6):
    ctx.set_option("pan_fast_stages", st); run(f"stages {st}")
ctx.set_option("pan_fast_stages", 4)
for r in (64, 128, 256, 512, 1024):
    ctx.set_option("pan_fast_rows", r); run(f"rows/tile {r}")
