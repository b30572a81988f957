"""host-buffer end-to-end timing of oip_pan_pipeline_host on C2 for several row-block sizes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from opticalimageprocessor_b200 import ops, synth
ctx = ops.Context(0, use_torch_stream=os.environ.get("OWN_STREAM") is None)
host_in = [torch.from_numpy(synth.strip_dn(bench.W, bench.ROWS, bench.SEED + i).byteswap()).pin_memory() for i in range(bench.N_CCD)]
kb_host = [torch.from_numpy(synth.rrc_coeffs(bench.W, bench.SEED + 100 + i)) for i in range(bench.N_CCD)]
out_w = ops.pan_out_width(bench.N_CCD, bench.W, bench.FOLD // 2)
host_out = torch.empty((bench.ROWS, out_w), dtype=torch.uint16).pin_memory()
px = bench.N_CCD * bench.W * bench.ROWS
for rows in [int(v) for v in os.environ.get("BLOCKS", "1024,2048,4096,8192").split(",")]:
    ctx.set_option("host_block_rows", rows)
    for _ in range(2):
        ops.pan_pipeline_host(ctx, host_in, kb_host, bench.DX, bench.DY, bench.FOLD // 2, host_out, fmt=ops.FMT_BE16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        ops.pan_pipeline_host(ctx, host_in, kb_host, bench.DX, bench.DY, bench.FOLD // 2, host_out, fmt=ops.FMT_BE16)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"block rows {rows}: {ms:.2f} ms  {px/ms/1e6:.2f} Gpx/s  ({(px*2)/ms/1e6:.1f} GB/s H2D, {host_out.numel()*2/ms/1e6:.1f} GB/s D2H)", flush=True)
