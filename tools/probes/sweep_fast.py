"""device-timed sweep of the fast PAN kernel's tunables on config C2 (development aid)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops, build
build.build()
ctx = ops.Context(0)
n, w, rows, f = 3, 8192, int(os.environ.get("ROWS", 32768)), 100
g = torch.Generator(device="cuda").manual_seed(1)
ccds = [torch.randint(64, 4032, (rows, w), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16) for _ in range(n)]
rng = np.random.default_rng(0)
kbs = []
for i in range(n):
    kb = np.empty((w, 2)); kb[:, 0] = 0.95 + 0.1 * rng.random(w); kb[:, 1] = 8 * rng.random(w)
    kbs.append(torch.from_numpy(kb).cuda())
dX, dY = [0, 1.37, -0.83], [0, -2.61, 3.19]
out = torch.empty((rows, ops.pan_out_width(n, w, f)), dtype=torch.uint16, device="cuda")
def run(K=10):
    for _ in range(3):
        ops.pan_pipeline(ctx, ccds, kbs, dX, dY, f, fmt=ops.FMT_BE16, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        ops.pan_pipeline(ctx, ccds, kbs, dX, dY, f, fmt=ops.FMT_BE16, out=out, check_error=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
px = n * w * rows
combos = os.environ.get("COMBOS")
if combos:
    combos = [tuple(int(v) for v in c.split(",")) for c in combos.split(";")]
else:
    combos = list(itertools.product([4, 3, 2], [3, 4, 6], [64, 128, 256]))
for minb, st, th in combos:
    ctx.set_option("pan_fast_minb", minb); ctx.set_option("pan_fast_stages", st); ctx.set_option("pan_fast_rows", th)
    ms = run()
    print(f"minb={minb} stages={st} rows={th}: {ms:.3f} ms  {px/ms/1e6:.1f} Gpx/s  {(px*2+out.numel()*2)/ms/1e6:.0f} GB/s", flush=True)
