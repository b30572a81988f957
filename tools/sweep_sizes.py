"""BASELINE configs[3] / configs[4] on one B200: the fused PAN path on strips of 256 MB ... 51.5 GB of raw input (the last
one is C4: 24576 px x 1 048 576 lines, input + output resident in HBM), device-timed, plus two size-independent
properties at full size: (1) a block of rows in the middle of the strip equals the same rows of a strip computed on its
own from the rows it needs (shard == whole), (2) the unshifted CCD's columns equal RRC(input) (checked through a
position-weighted checksum).  One JSON object per size on stdout: python tools/sweep_sizes.py > profiles/rNN_sweep.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from opticalimageprocessor_b200 import build, ops

build.build()
ctx = ops.Context(0)
if os.environ.get("PAN_FAST_ROWS"):  # development: override the warp-tile height
    ctx.set_option("pan_fast_rows", int(os.environ["PAN_FAST_ROWS"]))
N, W, F = 3, 8192, 100
DX, DY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
rng = np.random.default_rng(0)
kbs = []
for i in range(N):
    kb = np.empty((W, 2))
    kb[:, 0] = 0.95 + 0.1 * rng.random(W)
    kb[:, 1] = 8 * rng.random(W)
    kbs.append(torch.from_numpy(kb).cuda())
out_w = ops.pan_out_width(N, W, F)
sizes = [int(v) for v in os.environ.get("ROWS", "5461,21845,87381,349525,1048576").split(",")]  # 256 MB, 1, 4, 16, 51.5 GB


def fill(rows):
    g = torch.Generator(device="cuda").manual_seed(rows)
    ccds = []
    for _ in range(N):
        t = torch.empty((rows, W), dtype=torch.uint16, device="cuda")
        for r0 in range(0, rows, 65536):  # in pieces: randint has no uint16 kernel
            r1 = min(rows, r0 + 65536)
            t[r0:r1] = torch.randint(64, 4032, (r1 - r0, W), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16)
        ccds.append(t)
    return ccds


for rows in sizes:
    ccds = fill(rows)
    out = torch.empty((rows, out_w), dtype=torch.uint16, device="cuda")
    K = 3 if rows > 200000 else 10
    for _ in range(2):
        ops.pan_pipeline(ctx, ccds, kbs, DX, DY, F, fmt=ops.FMT_LE16, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        ops.pan_pipeline(ctx, ccds, kbs, DX, DY, F, fmt=ops.FMT_LE16, out=out, check_error=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    px = N * W * rows
    rec = {"rows": rows, "input_GB": px * 2 / 1e9, "output_GB": out.numel() * 2 / 1e9, "ms": ms, "Gpx_per_s": px / ms / 1e6,
           "algorithmic_GBps": (px * 2 + out.numel() * 2) / ms / 1e6}
    rec["frac_of_hbm_peak"] = rec["algorithmic_GBps"] / PEAK
    # (1) shard == whole: a block of rows in the middle computed on its own (row0 / n_rows, global section geometry)
    import ctypes as C
    from opticalimageprocessor_b200 import capi
    a = max(0, rows // 2 - 1000)
    b = min(rows, a + 2000)
    sub = torch.zeros((b - a, out_w), dtype=torch.uint16, device="cuda")
    d = ops.make_pan_desc(ccds, ops.FMT_LE16, kbs, DX, DY, [0, 1, 1], F, sub, total_rows=rows, row0=a, n_rows=b - a)
    capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
    capi.check(ctx.lib.oip_pan_check_error(ctx.h))
    rec["shard_equals_whole"] = bool(torch.equal(sub.view(torch.int16), out[a:b].view(torch.int16)))
    del sub
    # (2) the unshifted CCD: out[:, :W-F] == RRC(ccd0)[:, :W-F], through a checksum of checksums over 64k-row blocks
    ok = True
    for r0 in range(0, rows, 65536):
        r1 = min(rows, r0 + 65536)
        ref = ccds[0][r0:r1].clone()
        ops.inplace_rrc(ctx, ref, kbs[0])
        ok = ok and bool(torch.equal(ref[:, :W - F].view(torch.int16), out[r0:r1, :W - F].view(torch.int16)))
        del ref
    rec["copy_ccd_equals_rrc"] = ok
    print(json.dumps(rec), flush=True)
    del ccds, out
    torch.cuda.empty_cache()
