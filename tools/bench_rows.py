"""Per-row measurements for SURVEY 8(a): every hot-path row on one B200 against its HBM roofline, with the reference's
CPU path timed beside it on a bounded sample (oracle/_ref = the reference's own compiled code where it exists, else the
C restatement).  One JSON object per row on stdout; `python tools/bench_rows.py > profiles/rNN_rows.jsonl`.

Not the headline benchmark (that is bench.py); this is the coverage table of DESIGN.md section 6."""
import ctypes as C
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from opticalimageprocessor_b200 import build, ops, synth

build.build()
oracle.build()
ctx = ops.Context(0)
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
QUICK = os.environ.get("QUICK") == "1"


def dev_time(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3


def cpu_time(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def emit(row, what, units, unit_name, bytes_alg, secs, cpu_units, cpu_secs, cpu_kind, cpu_what):
    rec = {"row": row, "what": what, "ms": secs * 1e3, "throughput": units / secs / 1e9, "unit": f"G{unit_name}/s",
           "algorithmic_GBps": bytes_alg / secs / 1e9, "hbm_peak_GBps": PEAK, "frac_of_hbm_peak": bytes_alg / secs / 1e9 / PEAK,
           "cpu": {"throughput": cpu_units / cpu_secs / 1e9, "unit": f"G{unit_name}/s", "kind": cpu_kind, "sample": cpu_what},
           "speedup_vs_cpu": (units / secs) / (cpu_units / cpu_secs)}
    print(json.dumps(rec), flush=True)


rng = np.random.default_rng(1)
W = 12288  # reference geometry (ref oipshared.h:28)

# ------------------------------------------------------------------------------------------------ stage 1
# a reference-geometry downlink: n_frames image frames (1536 x 256 tiles) -> IMTR -> AOS, with anomalies
n_frames = 2 if QUICK else 6
imdt_np, _ = synth.make_imdt(n_frames, 1536, 256, seed=5)
imtr_np = synth.imtr_frames(imdt_np, chid=0x11)
aos_np = synth.aos_frames(imtr_np.reshape(-1))
file_np = synth.build_aos_file(aos_np, empty_every=64, bad_crc_at=set(range(100, aos_np.shape[0], 1024)))
buf = torch.from_numpy(file_np).cuda()
nb = buf.numel()
state = {}


def run_aos():
    state["off"], state["cnt"] = ops.aos_scan(ctx, buf)


t = dev_time(run_aos, it=5)
sample = file_np[: 16 << 20]
tc = cpu_time(lambda: oracle.aos_scan(sample))
emit("S1b-e", "oip_aos_scan: sync search + ValidateAosFrame + CRC-16 + chained skip rules (ref aux_separator.h:395-467,622-690)",
     nb, "B", nb * 1.0, t, sample.size, tc, "port", "C restatement (bit-wise CRC like CRC.h), 1 thread, 16 MiB of the same file")


def run_imtr():
    state["imdt"], state["st"] = ops.imtr_deframe(ctx, buf, state["off"])


t = dev_time(run_imtr, it=5)
n_pay = int(state["off"].numel())
off_np = state["off"].cpu().numpy().astype(np.uint64)
n_s = min(n_pay, 20000)
tc = cpu_time(lambda: oracle.imtr_deframe(file_np, off_np[:n_s]))
emit("S1f", "oip_imtr_deframe: 882-byte cadence, signature/type/CRC checks, 866-byte bodies (ref aux_separator.h:469-590)",
     n_pay * 880, "B", n_pay * (880 + 866.0), t, n_s * 880, tc, "port", f"C restatement, 1 thread, {n_s} payloads")

imdt = state["imdt"]


def run_index():
    state["ents"], state["fst"] = ops.image_frames_index(ctx, imdt, 1536, 256)


t = dev_time(run_index, it=5)
tc = cpu_time(lambda: oracle.image_frames(imdt_np, 1536, 256))
emit("S1g", "oip_image_frames_index: trailer signature search + chain / gap rules (ref aux_separator.h:627-656,287-320)",
     imdt.numel(), "B", imdt.numel() * 1.0, t, imdt_np.size, tc, "port", "C restatement: frame index AND tile unpack (one function), 1 thread, same IMDT")
nf = int(state["fst"][1])


def run_unpack():
    state["aux"], state["pan"], state["mss"] = ops.unpack_frames(ctx, imdt, 1536, 256, state["ents"], nf)


t = dev_time(run_unpack, it=5)
px = nf * 1280 * W
# the reference's own AuxSeparator on the IMDT file (= S1g + S1h + S1i incl. its file I/O on tmpfs)
cpu_kind, tc, cpu_px = "port", None, px
L = oracle.ref_oip_lib()
if L is not None:
    try:
        L.ref_auxsep.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_auxsep.restype = C.c_int
        d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        p = os.path.join(d, "KEL_MN200_CMOS-1_20220316_120309.IMDT")
        imdt_np.tofile(p)
        tc = cpu_time(lambda: L.ref_auxsep(p.encode(), d.encode()))
        cpu_kind = "reference"
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        os.rmdir(d)
    except Exception:
        tc = None
if tc is None:
    ents_np = oracle.image_frames(imdt_np, 1536, 256)
    tc = cpu_time(lambda: oracle.image_frames(imdt_np, 1536, 256))
emit("S1h-i", "oip_unpack_frames: aux copy + 40-tile de-interleave + BE->LE swap (ref aux_separator.h:335-393)",
     px, "px", px * 4.0 + nf * 49152 * 2, t, cpu_px, tc, cpu_kind,
     "AuxSeparator::SeparateImageData of the reference itself (oracle/_ref) on the same IMDT in /dev/shm, incl. its fwrite"
     if cpu_kind == "reference" else "C restatement of the frame index only")
del buf, imdt
state.clear()
torch.cuda.empty_cache()

# ------------------------------------------------------------------------------------------------ stage 2
rows = 8192 if QUICK else 32768
img = torch.randint(0, 4096, (rows, W), device="cuda", dtype=torch.int32).to(torch.uint16)
kb_np = synth.rrc_coeffs(W, 7)
kb = torch.from_numpy(kb_np).cuda()
t = dev_time(lambda: ops.inplace_rrc(ctx, img, kb))
smp = rng.integers(0, 4096, (2048, W), dtype=np.uint16)
if oracle.ref_oip_lib() is not None:
    tc, kind = cpu_time(lambda: oracle.ref_inplace_rrc(smp, kb_np)), "reference"
else:
    tc, kind = cpu_time(lambda: oracle.rrc(smp, kb_np)), "port"
emit("S2b-c", "oip_rrc_u16: IMO::InplaceRRC, fp64 k*s+b truncated (ref imageop.h:129-138)", rows * W, "px", rows * W * 4.0, t,
     smp.size, tc, kind, "IMO::InplaceRRC compiled from the reference's imageop.h, 1 thread (as the reference), 2048 lines")

# ------------------------------------------------------------------------------------------------ stage 3 (PAN)
src = img
import cv2
t = dev_time(lambda: ops.prestitch_shift(ctx, src, 1.37, -2.61))
crow = 4096
smp = rng.integers(0, 4096, (crow, W), dtype=np.uint16)
mx = (np.arange(W)[None, :] + np.zeros((crow, 1)) + 1.37).astype(np.float32)
my = (np.arange(crow)[:, None] + np.zeros((1, W)) - 2.61).astype(np.float32)
tc = cpu_time(lambda: cv2.remap(smp, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT))
emit("S3a", "oip_shift_cubic_u16: Stitcher::PreStitch + SectionaryRemap + cv::remap INTER_CUBIC (ref stitcher.h:83-139, imageop.h:230-275)",
     rows * W, "px", rows * W * 4.0, t, smp.size, tc, "reference",
     f"cv2.remap INTER_CUBIC (the reference's library call), {cv2.getNumThreads()} OpenCV threads, {crow} lines, map build not counted")

img2 = torch.randint(0, 4096, (rows, W), device="cuda", dtype=torch.int32).to(torch.uint16)
t = dev_time(lambda: ops.stitch_big_raw(ctx, [src, img2], 100))
a, b = smp, rng.integers(0, 4096, (crow, W), dtype=np.uint16)
tc = cpu_time(lambda: oracle.stitch_concat([a, b], 100))
emit("S3b-c", "oip_stitch_concat_u16: IMO::StitchBigRaw hard cut L[0:W-f] | R[f:W] (ref imageop.h:277-363)", 2 * rows * W, "px",
     2 * rows * (W - 100) * 4.0, t, 2 * crow * W, tc, "port", f"C restatement (memcpy per line), 1 thread, {crow} lines, in memory")

kb2 = torch.from_numpy(synth.rrc_coeffs(W, 8)).cuda()
out = torch.empty((rows, ops.pan_out_width(2, W, 100)), dtype=torch.uint16, device="cuda")
be = [src.view(torch.int16).clone().view(torch.uint16), img2]
t = dev_time(lambda: ops.pan_pipeline(ctx, be, [kb, kb2], [0, 1.37], [0, -2.61], 100, out=out, check_error=False))


def cpu_fused():
    r0 = oracle.ref_inplace_rrc(a, kb_np) if oracle.ref_oip_lib() is not None else oracle.rrc(a, kb_np)
    r1 = oracle.ref_inplace_rrc(b, kb_np) if oracle.ref_oip_lib() is not None else oracle.rrc(b, kb_np)
    r1 = cv2.remap(r1, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    return np.concatenate([r0[:, :W - 100], r1[:, 100:]], axis=1)


tc = cpu_time(cpu_fused)
emit("S2+S3 fused, reference geometry", "oip_pan_pipeline: 2 CMOS x 12288 px, RRC + shift of CMOS-2 + stitch in one pass", 2 * rows * W, "px",
     2 * rows * W * 2.0 + out.numel() * 2.0, t, 2 * crow * W, tc, "reference" if oracle.ref_oip_lib() is not None else "port",
     f"InplaceRRC x2 (1 thread) + cv2.remap ({cv2.getNumThreads()} threads) + concat, {crow} lines, in memory")
del img2, out, be
torch.cuda.empty_cache()

# ------------------------------------------------------------------------------------------------ stage 3 (MSS)
lines = 4096 if QUICK else 16384
wb = W // 4
mss = torch.randint(0, 4096, (lines, W), device="cuda", dtype=torch.int32).to(torch.uint16)
kbs_np = [synth.rrc_coeffs(wb, 20 + i) for i in range(4)]
kbs = [torch.from_numpy(k).cuda() for k in kbs_np]
cX = [[0.8 + 0.1 * i, -1.5e-4 * (i + 1)] for i in range(4)]
cY = [[-3.2 + i, 2e-4 * (i + 1), -1e-8 * (i - 1.5)] for i in range(4)]
res = {}


def run_mss():
    res["n"], res["out"] = ops.band_align(ctx, mss, wb, kbs, cX, cY)


t = dev_time(run_mss, it=5)
n_out = res["n"]
cl = 2048
msmp = rng.integers(0, 4096, (cl, W), dtype=np.uint16)


def cpu_mss():
    planes = [oracle.rrc(p, k) for p, k in zip(oracle.mss_split(msmp), kbs_np)]
    return oracle.band_align(planes, cX, cY, min_process_lines=100)


tc = cpu_time(cpu_mss)
emit("S2d+S3d", "oip_band_align_merge: band split + RRC x4 + per-band polynomial cubic remap + 4-channel merge (ref preproc.h:56-80,202-222,351-468)",
     lines * W, "px", lines * W * 2.0 + n_out * wb * 8.0, t, cl * W, tc, "port",
     f"C restatement (RRC + remap restated from cv2, pinned bit-exact) 1 thread, {cl} lines")

c4a = res["out"][: (n_out // 2) * 2].reshape(-1, wb, 4)
c4b = c4a.clone()
t = dev_time(lambda: ops.stitch_tiff_geometry(ctx, [c4a, c4b], 25, band_map=[3, 2, 1, 4]))
ca = rng.integers(0, 4096, (2048, wb, 4), dtype=np.uint16)
tc = cpu_time(lambda: oracle.stitch_concat_c4([ca, ca], 25, [3, 2, 1, 4]))
emit("S3e", "oip_stitch_concat_c4: StitchTiff* geometry on CV_16UC4 + band map (ref imageop.h:416-421,501-538)", 2 * c4a.shape[0] * wb * 4, "sample",
     2 * c4a.shape[0] * (wb - 25) * 8 * 2.0, t, 2 * ca.size, tc, "port", "C restatement, 1 thread, 2048 lines, in memory")

# ------------------------------------------------------------------------------------------------ extension
packed = torch.from_numpy(synth.pack_bits(rng.integers(0, 4096, (4096, W), dtype=np.uint16), 12)).cuda()
t = dev_time(lambda: ops.unpack_lines(ctx, packed, ops.FMT_PACK12, W))
psm = packed[:1024].cpu().numpy()
tc = cpu_time(lambda: oracle.unpack_bits(psm, 12, W, 1024, psm.shape[1]))
emit("ext", "oip_unpack_lines: MSB-first packed 12-bit -> u16 (not in the reference, SURVEY 0.1)", 4096 * W, "px", 4096 * W * 3.5, t,
     1024 * W, tc, "port", "C restatement, 1 thread, 1024 lines")

# ------------------------------------------------------------------------------------------------ N1
# Stitcher::CalcSttParameters at the reference defaults: 10 sections x 16000 lines x 200 overlap columns of two
# 12288-px strips.  Algorithmic bytes: the two u16 slices, read once.  CPU: the same loop on cv2.phaseCorrelate.
import cv2  # noqa: E402

st_lines = 10 * 16000 + 11 * 400
p1 = torch.from_numpy(rng.integers(64, 4032, (st_lines, W), dtype=np.uint16)).cuda()
p2 = p1.view(torch.int16).roll(shifts=(3, 0), dims=(0, 1)).contiguous().view(torch.uint16)
p2[:, :200] = p1[:, W - 200:].view(torch.int16).roll(shifts=(2, 1), dims=(0, 1)).view(torch.uint16)
t = dev_time(lambda: ops.calc_stt_parameters(ctx, p1, p2), warm=2, it=5)
s1 = p1[400:16400, W - 200:].cpu().numpy().astype(np.float32)
s2 = p2[400:16400, :200].cpu().numpy().astype(np.float32)
tc = cpu_time(lambda: cv2.phaseCorrelate(s1, s2), reps=3)
emit("N1", "oip_stt_parameters: CalcSttParameters = 10 x phase correlation of 16000 x 200 overlap slices, cuFFT + own kernels (ref stitcher.h:148-201)",
     10 * 16000 * 200 * 2, "px", 10 * 16000 * 200 * 2 * 2.0, t, 16000 * 200 * 2, tc, "reference",
     f"cv2.phaseCorrelate ({cv2.getNumThreads()} OpenCV threads), one 16000 x 200 section")

# ------------------------------------------------------------------------------------------------ fused C2, three raw sample formats
# north_star: unpack -> correct -> stitch in one pass.  3 CCD x 8192 px x 32768 lines; algorithmic bytes per input px =
# b_in + 2 * W_out / W_in with b_in = 2 (BE16), 1.5 (12-bit packed), 1.25 (10-bit packed)  
del p1, p2
torch.cuda.empty_cache()
Wc, Rc, fc = 8192, (8192 if QUICK else 32768), 100
dXc, dYc = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
kbc = [torch.from_numpy(synth.rrc_coeffs(Wc, 300 + i)).cuda() for i in range(3)]
outc = torch.empty((Rc, ops.pan_out_width(3, Wc, fc)), dtype=torch.uint16, device="cuda")
px_c = 3 * Wc * Rc
for name, fmt, b_in, bits in [("BE16", ops.FMT_BE16, 2.0, 16), ("12-bit packed", ops.FMT_PACK12, 1.5, 12), ("10-bit packed", ops.FMT_PACK10, 1.25, 10)]:
    ccds = []
    for i in range(3):
        dn = synth.strip_dn(Wc, Rc, 400 + i)
        if bits == 10:
            dn = dn >> 2
        ccds.append(torch.from_numpy(dn.byteswap() if bits == 16 else synth.pack_bits(dn, bits)).cuda())
    t = dev_time(lambda: ops.pan_pipeline(ctx, ccds, kbc, dXc, dYc, fc, fmt=fmt, out=outc, check_error=False, w=Wc))
    emit(f"fused C2 {name}", f"oip_pan_pipeline: 3 CCD x {Wc} px x {Rc} lines, {name} samples in, RRC + cubic shift of 2 CCDs + stitch, one pass",
         px_c, "px", px_c * b_in + outc.numel() * 2.0, t, 1, 1.0, "n/a", "see the BE16 row of bench.py for the CPU reference")
    del ccds
