"""does the number of distinct code variants resident on an SM matter? (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops, build
build.build()
ctx = ops.Context(0)
ctx.set_option("pan_fast_minb", int(os.environ.get("MINB", 3)))
w, rows, f = 8192, 32768, 100
g = torch.Generator(device="cuda").manual_seed(1)
ccds = [torch.randint(64, 4032, (rows, w), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16) for _ in range(3)]
rng = np.random.default_rng(0)
kbs = []
for i in range(3):
    kb = np.empty((w, 2)); kb[:, 0] = 0.95 + 0.1 * rng.random(w); kb[:, 1] = 8 * rng.random(w)
    kbs.append(torch.from_numpy(kb).cuda())
def run(n, dX, dY, shifted, fmt, K=10, label=""):
    out = torch.empty((rows, ops.pan_out_width(n, w, f)), dtype=torch.uint16, device="cuda")
    for _ in range(3):
        ops.pan_pipeline(ctx, ccds[:n], kbs[:n], dX, dY, f, fmt=fmt, out=out, shifted=shifted)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        ops.pan_pipeline(ctx, ccds[:n], kbs[:n], dX, dY, f, fmt=fmt, out=out, shifted=shifted, check_error=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    nrem = sum(shifted)
    print(f"{label:52s} {ms:.3f} ms   {n*w*rows/ms/1e6:7.1f} Gpx/s   per REMAP CCD {ms/max(nrem,1):.3f} ms", flush=True)
BE, LE = ops.FMT_BE16, ops.FMT_LE16
run(3, [0, 1.37, -0.83], [0, -2.61, 3.19], [0, 1, 1], BE, label="C2: copy + 2 remap variants (DM differ), BE")
run(3, [0, 1.37, 1.37], [0, -2.61, -2.61], [0, 1, 1], BE, label="copy + 2 remap, same variant, BE")
run(3, [1.37, 1.37, 1.37], [-2.61, -2.61, -2.61], [1, 1, 1], BE, label="3 remap, same variant, BE")
run(3, [0, 0, 0], [0, 0, 0], [0, 0, 0], BE, label="3 copy, BE")
run(1, [1.37], [-2.61], [1], BE, label="1 remap only, BE")
run(1, [1.37], [-2.61], [1], LE, label="1 remap only, LE")
run(1, [0.37], [-2.61], [1], LE, label="1 remap only, LE, DM even")
run(2, [0, 1.37], [0, -2.61], [0, 1], BE, label="copy + 1 remap, BE")
