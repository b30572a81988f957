import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_framed as b
from opticalimageprocessor_b200 import capi, ops, synth
import oracle
ctx = ops.Context(0)
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
file_np, pan = b.make_downlink(lines)
buf = torch.from_numpy(file_np).cuda()
kb_np = [synth.rrc_coeffs(b.W, b.SEED + 100 + i) for i in range(3)]
kbs = [torch.from_numpy(k).cuda() for k in kb_np]
off, cnt = ops.aos_scan(ctx, buf)
o_off, o_cnt = oracle.aos_scan(file_np)
print("aos", cnt.tolist(), o_cnt.tolist(), np.array_equal(off.cpu().numpy().astype(np.uint64), o_off))
imdt, ist = ops.imtr_deframe(ctx, buf, off)
o_imdt, o_st = oracle.imtr_deframe(file_np, o_off)
print("imtr", ist.tolist(), o_st.tolist(), imdt.numel(), o_imdt.size, np.array_equal(imdt.cpu().numpy(), o_imdt))
ents, fst = ops.image_frames_index(ctx, imdt, b.TC, b.TL)
n, _, o_pan, _, o_fst = oracle.image_frames(o_imdt, b.TC, b.TL)
print("frames", fst.tolist(), o_fst.tolist(), n, np.array_equal(o_pan, pan))
tab = ops.frame_tile_table(ents, int(fst[1]))
print("tab mod 16:", sorted(set((tab[:, 0] % 16).tolist())), "mod 4:", sorted(set((tab.reshape(-1) % 4).tolist())))
keep = []
out, desc = ops.pan_pipeline_from_frames(ctx, [imdt] * 3, [tab] * 3, b.TC, b.TL, kbs, b.DX, b.DY, b.FOLD // 2, keep=keep)
st = (C.c_int64 * 4)()
capi.check(ctx.lib.oip_pan_plan_coverage(C.byref(desc), 1, 128, None, st))
print("coverage generic px, fast px, generic tiles, fast tiles:", list(st))
for name, fn in [("pan tiles", lambda: ops.pan_pipeline_from_frames(ctx, [imdt] * 3, [tab] * 3, b.TC, b.TL, kbs, b.DX, b.DY, b.FOLD // 2, out=out, keep=keep, check_error=False)),
                 ("imtr", lambda: ops.imtr_deframe(ctx, buf, off)), ("aos", lambda: ops.aos_scan(ctx, buf))]:
    fn(); torch.cuda.synchronize()
    for rep in range(3):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); fn(); e.record(); torch.cuda.synchronize()
        print(name, "event ms", a.elapsed_time(e), "wall ms", (time.perf_counter() - t0) * 1e3)
rows = np.array(sorted(set(range(16)) | set(range(1016, 1032)) | set(range(lines - 16, lines))), np.int64)
want = oracle.pan_rows(lambda i, a, bb: pan[a:bb], 3, b.W, kb_np, b.DX, b.DY, b.FOLD // 2, lines, rows)
g = out.view(torch.int16)[torch.from_numpy(rows).cuda()].cpu().numpy().view(np.uint16)
bad = np.argwhere(g != want)
print("pan rows mismatches", len(bad), bad[:8].tolist(), "rows", sorted(set(rows[bad[:, 0]].tolist()))[:20] if len(bad) else "")
