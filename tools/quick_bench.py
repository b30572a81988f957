"""quick device-timed check of the fused PAN kernel on config C2 (development aid, not bench.py)."""
import sys, os, time
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
for fmt, name in [(ops.FMT_LE16, "LE16"), (ops.FMT_BE16, "BE16")]:
    for _ in range(3):
        ops.pan_pipeline(ctx, ccds, kbs, dX, dY, f, fmt=fmt, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        ops.pan_pipeline(ctx, ccds, kbs, dX, dY, f, fmt=fmt, out=out, check_error=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    px = n * w * rows
    byts = px * 2 + out.numel() * 2
    print(f"{name}: {ms:.3f} ms/step  {px/ms/1e6:.1f} Gpx/s  {byts/ms/1e6:.1f} GB/s algorithmic")
# stand-alone RRC
img = ccds[0].clone()
for _ in range(3): ops.inplace_rrc(ctx, img, kbs[0])
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.inplace_rrc(ctx, img, kbs[0])
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
print(f"RRC alone: {ms:.3f} ms  {img.numel()/ms/1e6:.1f} Gpx/s  {img.numel()*4/ms/1e6:.1f} GB/s")
# copy baseline
a = torch.empty(512 << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); e0.record()
for _ in range(10): b.copy_(a)
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
print(f"torch copy 512MiB: {2*a.numel()/ms/1e6:.1f} GB/s")
