"""ncu driver: REMAP-only launch of pan_fast_kernel (one shifted CCD 8192 x 32768, BE16 + RRC)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops
ctx = ops.Context(0)
ctx.set_option("pan_fast_minb", int(os.environ.get("MINB", 3)))
w, rows = 8192, 32768
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.randint(64, 4032, (rows, w), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16)
rng = np.random.default_rng(0)
kb = np.empty((w, 2)); kb[:, 0] = 0.95 + 0.1 * rng.random(w); kb[:, 1] = 8 * rng.random(w)
kb = torch.from_numpy(kb).cuda()
out = torch.empty((rows, w), dtype=torch.uint16, device="cuda")
for _ in range(5):
    ops.pan_pipeline(ctx, [src], [kb], [1.37], [-2.61], 0, fmt=ops.FMT_BE16, out=out, shifted=[1], check_error=False)
torch.cuda.synchronize()
print("done")
