"""C2 through the fast kernel in its input variants (BE16 lines, packed 12-bit, frame tiles): ms per pass"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops, synth
ctx = ops.Context(0)
W, R, f = 8192, 32768, 100
dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
kb = [torch.from_numpy(synth.rrc_coeffs(W, 300 + i)).cuda() for i in range(3)]
out = torch.empty((R, ops.pan_out_width(3, W, f)), dtype=torch.uint16, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
dn = [synth.strip_dn(W, R, 400 + i) for i in range(3)]
be = [torch.from_numpy(d.byteswap()).cuda() for d in dn]
print("BE16 lines   %.3f ms" % t(lambda: ops.pan_pipeline(ctx, be, kb, dX, dY, f, fmt=ops.FMT_BE16, out=out, check_error=False)))
del be
p12 = [torch.from_numpy(synth.pack_bits(d, 12)).cuda() for d in dn]
print("packed 12    %.3f ms" % t(lambda: ops.pan_pipeline(ctx, p12, kb, dX, dY, f, fmt=ops.FMT_PACK12, out=out, check_error=False, w=W)))
del p12
import bench_framed as bf
file_np, pan = bf.make_downlink(R)
buf = torch.from_numpy(file_np).cuda()
off, cnt = ops.aos_scan(ctx, buf); imdt, st = ops.imtr_deframe(ctx, buf, off); ents, fst = ops.image_frames_index(ctx, imdt, bf.TC, bf.TL)
tab = ops.frame_tile_table(ents, int(fst[1])); keep = []
print("frame tiles  %.3f ms" % t(lambda: ops.pan_pipeline_from_frames(ctx, [imdt] * 3, [tab] * 3, bf.TC, bf.TL, kb, dX, dY, f, out=out, keep=keep, check_error=False)))
