"""REMAP-only timing (one shifted CCD, 8192 x 32768) -- development aid for kernel variants (OIP_B200_LIB=...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops
ctx = ops.Context(0)
w, rows = 8192, 32768
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.randint(64, 4032, (rows, w), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16)
rng = np.random.default_rng(0)
kb = np.empty((w, 2)); kb[:, 0] = 0.95 + 0.1 * rng.random(w); kb[:, 1] = 8 * rng.random(w)
kb = torch.from_numpy(kb).cuda()
out = torch.empty((rows, w), dtype=torch.uint16, device="cuda")
for minb in (3, 4):
    ctx.set_option("pan_fast_minb", minb)
    for dX in (1.37, 0.37):
        for use_kb in (True, False):
            args = dict(fmt=ops.FMT_BE16, out=out, shifted=[1], check_error=False)
            for _ in range(3): ops.pan_pipeline(ctx, [src], [kb] if use_kb else None, [dX], [-2.61], 0, **args)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ops.pan_pipeline(ctx, [src], [kb] if use_kb else None, [dX], [-2.61], 0, **args)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            cyc = ms * 1e-3 * 1.965e9 * 592 / (rows * w / 248)
            print(f"minb={minb} dX={dX} rrc={int(use_kb)}: {ms:.3f} ms  {w*rows/ms/1e6:6.1f} Gpx/s  ~{cyc:.0f} SMSP cycles per warp-row", flush=True)
