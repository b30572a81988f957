"""ncu / timing driver for stage 1 (frame handling) on a reference-geometry downlink of N image frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opticalimageprocessor_b200 import ops, synth
ctx = ops.Context(0)
n_frames = int(os.environ.get("FRAMES", 6))
rep = int(os.environ.get("REP", 1))
imdt_np, _ = synth.make_imdt(n_frames, 1536, 256, seed=5)
imtr_np = synth.imtr_frames(imdt_np, chid=0x11)
aos_np = synth.aos_frames(imtr_np.reshape(-1))
file_np = synth.build_aos_file(aos_np, empty_every=64, bad_crc_at=set(range(100, aos_np.shape[0], 1024)))
if rep > 1:
    file_np = np.tile(file_np, rep)   # IMTR sequence numbers repeat: only warnings (ref aux_separator.h:530-533)
buf = torch.from_numpy(file_np).cuda()
print("file bytes", buf.numel())
def ev():
    return torch.cuda.Event(enable_timing=True)
variants = [int(v) for v in os.environ.get("IMTR_RUNS", "1").split(",")]   # 1: run-based gather (default), 0: per-frame gather; "0,1" = A/B
for runs, it in [(r, i) for r in variants for i in range(5)]:
    if it == 0:
        ctx.set_option("imtr_runs", runs)
        print("imtr_runs =", runs)
    t = [ev() for _ in range(5)]
    t[0].record()
    off, cnt = ops.aos_scan(ctx, buf); t[1].record()
    imdt, st = ops.imtr_deframe(ctx, buf, off); t[2].record()
    ents, fst = ops.image_frames_index(ctx, imdt, 1536, 256); t[3].record()
    aux, pan, mss = ops.unpack_frames(ctx, imdt, 1536, 256, ents, int(fst[1])); t[4].record()
    torch.cuda.synchronize()
    ms = [t[i].elapsed_time(t[i + 1]) for i in range(4)]
    print("aos_scan %.3f ms (%.0f GB/s)  imtr_deframe %.3f ms (%.0f GB/s of payload)  index %.3f ms (%.0f GB/s)  unpack %.3f ms (%.0f GB/s r+w)" % (
        ms[0], buf.numel() / ms[0] / 1e6, ms[1], off.numel() * 880 / ms[1] / 1e6, ms[2], imdt.numel() / ms[2] / 1e6, ms[3],
    (pan.numel() + mss.numel()) * 4 / ms[3] / 1e6), "frames", int(fst[1]), "counters", cnt.tolist())
