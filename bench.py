#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused PAN hot path (unpack -> RRC -> shift -> stitch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[3], the configuration north_star quotes its target on: a 24576 px (3 CCD x 8192)
x 1 048 576-line panchromatic strip, 16-bit big-endian raw samples (the byte order inside the downlink sub-images), fold 200,
seam offsets (1.37, -2.61) / (-0.83, 3.19).  It fits one B200 (51.5 GB in + 50.7 GB out), so N = 1 runs the whole strip and
N > 1 splits the SAME strip into N scanline blocks (strong scaling).  Section geometry is a function of the global line
index; the few halo rows a block needs from its neighbours are read by the kernel straight from the neighbour GPU's
memory over NVLink (CUDA IPC peer mappings), no data-path collective.  One step = one pass of the fused kernels over
the rank's block.

value    : Gpixel/s, device-timed (CUDA events on the launching stream), inputs resident in HBM, whole job over all
           ranks, max over ranks.  Warm-up runs until the GPU has worked for >= 2 s (so the clocks are the sustained
           ones), never fewer than --warmup / 3 steps.
parity   : outside the timed region every rank compares a bounded set of its output rows (first / last 16 rows of its
           block, +-8 rows around every global section edge of both shifted CCDs, the stale bottom rows) with the CPU
           oracle (oracle.pan_rows) -> "parity_ok".
e2e      : same metric through the host-buffer C-ABI call (oip_pan_pipeline_host: pinned host in/out, H2D + kernels +
           D2H inside the timed call), every rank on its own block, max over ranks.
roofline : algorithmic HBM bytes of the fused kernel / its mean launch time vs the measured copy peak.
cpu_baseline / --impl reference : the reference's CPU path for the same work = fp64 RRC loop (single thread, ref
           imageop.h:129-138) + cv2.remap in 30000-row sections with OpenCV's own thread pool (the library call at
           imageop.h:258) + line concat, on a bounded slice of the same strip.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CCD, W, TOTAL_ROWS, FOLD = 3, 8192, 1 << 20, 200
DX = [0.0, 1.37, -0.83]
DY = [0.0, -2.61, 3.19]
SEED = 0x0A11CE04
METRIC = "Gpixel/s (raw->stitched, device-timed)"
WORKLOAD = ("C4: 24576 px (3 CCD x 8192) x 1048576 lines, 16-bit BE raw in, fold 200, RRC + cubic shift + stitch; "
            "the strip is split into N scanline blocks (strong scaling)")
WARM_SECONDS = float(os.environ.get("OIP_BENCH_WARM_SECONDS", "2.0"))   # (0 under ncu: a launch list needs no 2 s of warm-up)
REF_ROWS = 59996   # lines one step of the CPU reference arm processes: exactly two full 30000-row sections for both shifted CCDs (advance 29997 / 29996), like the 35 sections of the whole strip


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def algorithmic_bytes_per_px():
    out_w = N_CCD * W - 2 * (N_CCD - 1) * (FOLD // 2)
    return 2.0 + 2.0 * out_w / (N_CCD * W)  # b_in + 2*W_out/W_in


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs (one sample is taken synchronously at
    start and at stop, so a region shorter than the sampling period still has data)"""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, dev):
        self.dev, self.samples, self.reasons, self.max_mhz, self.h, self.nv = dev, [], set(), None, None, None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        if self.h is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, nm in self.NAMES.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.005)

    def start(self):
        self._t.start()

    def stop(self):
        self._sample()  # still under load: the caller stops the sampler before it synchronises
        self._stop.set()
        self._t.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_pass(rows: int, threads=None):
    """the reference's CPU path on the first `rows` lines of the strip (as a strip of its own); returns seconds (in memory,
    no disk; the map fill of stitcher.h:93-99 is left out of the timed region -- generous to the reference)"""
    import cv2
    import oracle
    from opticalimageprocessor_b200 import synth
    oracle.build()
    if threads:
        cv2.setNumThreads(threads)
    S, G = 30000, 32767                                                   # ref imageop.h:19-20
    ccds = [synth.strip_dn(W, rows, SEED + i) for i in range(N_CCD)]
    kbs = [synth.rrc_coeffs(W, SEED + 100 + i) for i in range(N_CCD)]
    hb = rows if rows <= G else S
    mx = [(np.arange(W)[None, :] + np.zeros((hb, 1)) + DX[i]).astype(np.float32) for i in range(N_CCD)]
    my = [(np.arange(hb)[:, None] + np.zeros((1, W)) + DY[i]).astype(np.float32) for i in range(N_CCD)]
    f = FOLD // 2
    rrc = oracle.ref_inplace_rrc if oracle.ref_oip_lib() is not None else oracle.rrc   # the reference's own compiled loop if built
    buff = np.zeros((S, W), np.uint16) if rows > G else None
    t0 = time.perf_counter()
    parts = []
    for i in range(N_CCD):
        r = rrc(ccds[i], kbs[i])                                          # ref imageop.h:129-138 (1 thread)
        if i > 0 and rows <= G:                                           # one cv::remap (the reference throws here, imageop.h:242-244)
            r = cv2.remap(r, mx[i], my[i], cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
        elif i > 0:                                                       # ref stitcher.h:83-139 / imageop.h:246-272
            ucut = 0 if DY[i] >= 0 else int(-DY[i]) + 1
            bcut = int(DY[i]) + 1 if DY[i] >= 0 else 0
            outp, off, sec, dst = [], 0, 0, None
            while True:
                n = min(S, rows - off)
                if n <= ucut + bcut:
                    break
                buff[:n] = r[off:off + n]
                dst = cv2.remap(buff, mx[i], my[i], cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)   # imageop.h:258
                if sec == 0 and ucut > 0:
                    outp.append(dst[:ucut])
                outp.append(dst[ucut:n - bcut])
                off += n - ucut - bcut
                sec += 1
            if bcut > 0:
                outp.append(dst[S - bcut:S])
            r = np.concatenate(outp)
        lo, hi = (0 if i == 0 else f), (W if i == N_CCD - 1 else W - f)
        parts.append(r[:, lo:hi])
    out = np.concatenate(parts, axis=1)                                   # ref imageop.h:340-355
    dt = time.perf_counter() - t0
    assert out.shape[0] == rows
    return dt, int(out[::97, ::89].astype(np.int64).sum())


def cpu_kind():
    """'reference' when the RRC loop is the reference's own imageop.h compiled under oracle/_ref (the resampling is the
    reference's library call, cv2.remap, either way); 'port' when only the C restatement is available"""
    try:
        import oracle
        return "reference" if oracle.ref_oip_lib() is not None else "port"
    except Exception:
        return "port"


def run_reference(args):
    """the reference's own CPU implementation of the path on the box's host cores; one step = REF_ROWS lines of the C4
    strip (a bounded sample: the whole strip would take ~1 minute per step)"""
    import cv2
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = REF_ROWS
    for _ in range(max(0, args.warmup)):
        cpu_reference_pass(2048)
    ts = []
    for _ in range(args.steps):
        dt, _ = cpu_reference_pass(rows)
        ts.append(dt)
    dt = float(np.mean(ts))
    px = N_CCD * W * rows
    val = px / dt / 1e9
    cores = cv2.getNumThreads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 RRC / f32 bicubic / u16", "data": "synthetic",
        "config": bench_config(int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": val, "unit": "Gpixel/s", "cores": cores, "kind": cpu_kind(),
                         "sample": f"{rows} lines (of the 1048576) of the same 3 x {W} px strip per step: IMO::InplaceRRC (the "
                                   f"reference's imageop.h compiled under oracle/_ref when present, else its C restatement; 1 thread "
                                   f"like the reference) + cv2.remap INTER_CUBIC in 30000-row sections ({cores} OpenCV threads, the "
                                   f"reference's own library call) + concat, in memory"},
        "e2e": {"value": val, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def bench_config(world: int, total_rows: int = TOTAL_ROWS) -> dict:
    return {"workload": WORKLOAD if total_rows == TOTAL_ROWS else WORKLOAD.replace("1048576 lines", f"{total_rows} lines (--rows)"),
            "n_ccd": N_CCD, "w": W, "total_rows": total_rows, "rows_per_gpu": total_rows // world,
            "fold_cols": FOLD, "dX": DX, "dY": DY,
            "l2": "inputs (51.5 GB / N) and output (50.7 GB / N) per step >> 126 MB L2",
            "multi_gpu": ("scanline-block shards of ONE strip, halo rows read from peer HBM over NVLink (CUDA IPC), "
                          "no data-path collective") if world > 1 else "single GPU, whole strip"}


# ------------------------------------------------------------------------------------------ GPU arm
def parity_rows(first: int, last: int, total_rows: int):
    """the bounded set of global output rows a rank checks against the oracle: first / last 16 rows of its block, +-8 rows
    around every run of rows SectionaryRemap writes (section edges s*(30000-ucut-bcut), ref imageop.h:260-272) for both
    shifted CCDs, and the last rows of the strip (stale rows of the partial last section)"""
    import oracle
    want = set(range(first, min(last, first + 16))) | set(range(max(first, last - 16), last))
    for i in range(1, N_CCD):
        pieces, _ = oracle.shift_pieces(total_rows, DY[i])
        for o0, n, *_ in pieces:
            for g in range(o0 - 8, o0 + 8):
                if first <= g < last:
                    want.add(g)
    for g in range(total_rows - 16, total_rows):
        if first <= g < last:
            want.add(g)
    return np.array(sorted(want), np.int64)


def run_gpu(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from opticalimageprocessor_b200 import build, capi, ops, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()
    ctx = ops.Context(local)
    lib = ctx.lib
    dev = torch.device("cuda", local)
    total_rows = args.rows
    first, last = sharding.shard_range(total_rows, world, rank)
    rows = last - first
    out_w = ops.pan_out_width(N_CCD, W, FOLD // 2)
    f = FOLD // 2

    def dev_alloc(nbytes):
        p = C.c_void_p()
        capi.check(lib.oip_dev_alloc(ctx.h, nbytes, C.byref(p)))
        return p.value

    # ---- the rank's block of the synthetic strip, generated on the device (BE16 = byte order inside the downlink
    #      sub-images); raw cudaMalloc allocations because a CUDA-IPC handle maps a whole allocation
    d_in = [dev_alloc(rows * W * 2) for _ in range(N_CCD)]
    for i in range(N_CCD):
        capi.check(lib.oip_synth_strip_dn(ctx.h, d_in[i], W, rows, first, W, SEED + i, 1))
    d_out = torch.empty((rows, out_w), dtype=torch.uint16, device=dev)
    d_out_ptr = d_out.data_ptr()
    ctx.sync()
    kb_np = [synth.rrc_coeffs(W, SEED + 100 + i) for i in range(N_CCD)]
    kb_host = [torch.from_numpy(k) for k in kb_np]
    d_kb = [t.to(dev) for t in kb_host]

    class Shape:  # make_pan_desc only reads .shape of the CCD arguments when explicit segments are given
        def __init__(self, r, c):
            self.shape = (r, c)

    class OutPtr:
        def __init__(self, ptr, pitch):
            self._p, self._s = ptr, pitch

        def data_ptr(self):
            return self._p

        def stride(self, k):
            return self._s

    desc = ops.make_pan_desc([Shape(rows, W)] * N_CCD, ops.FMT_BE16, d_kb, DX, DY, [i > 0 for i in range(N_CCD)], f,
                             d_out, total_rows=total_rows, row0=first, n_rows=rows,
                             segs=[[(d_in[i], first, rows, W * 2)] for i in range(N_CCD)])
    opened = []
    if world > 1:  # halo rows from the neighbour blocks: peer mappings over NVLink (no collective on the data path)
        handles = []
        for i in range(N_CCD):
            hb = C.create_string_buffer(64)
            capi.check(lib.oip_ipc_export(ctx.h, C.c_void_p(d_in[i]), hb))
            handles.append(hb.raw)
        allh = [None] * world
        dist.all_gather_object(allh, handles)
        peer_ptr = {}

        def peer(r, i):
            if (r, i) not in peer_ptr:
                p = C.c_void_p()
                capi.check(lib.oip_ipc_open(ctx.h, allh[r][i], C.byref(p)))
                peer_ptr[(r, i)] = p.value
                opened.append(p.value)
            return peer_ptr[(r, i)]

        sharding.attach_segments(desc, N_CCD, total_rows, world, rank, d_in, W * 2, peer)
        dist.barrier()

    def step():
        capi.check(lib.oip_pan_pipeline(ctx.h, C.byref(desc)))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up: >= W (>= 3) steps AND >= WARM_SECONDS of GPU work, so the timed steps run at the sustained clocks
    n_warm = 0
    t_w0 = time.perf_counter()
    while True:
        step()
        n_warm += 1
        if n_warm >= max(args.warmup, 3):
            torch.cuda.synchronize()
            busy = torch.tensor([time.perf_counter() - t_w0], device=dev)
            if world > 1:
                dist.all_reduce(busy, op=dist.ReduceOp.MIN)  # every rank leaves the loop at the same step
            if busy.item() >= WARM_SECONDS or n_warm >= 4000:
                break
    capi.check(lib.oip_pan_check_error(ctx.h))
    sync_all()
    l0 = ctx.launches
    clocks = ClockSampler(local)
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    clk = clocks.stop()   # one more sample while the queued steps are still running
    sync_all()
    launches = ctx.launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]))
    capi.check(lib.oip_pan_check_error(ctx.h))

    # ---- parity (outside the timed region): a bounded set of this rank's output rows against the CPU oracle
    parity_ok, parity_n, parity_bad = True, 0, 0
    if not args.no_parity:
        import oracle
        oracle.build()
        chk = parity_rows(first, last, total_rows)
        got = np.empty((chk.size, out_w), np.uint16)
        row_b = out_w * 2
        for k, g in enumerate(chk):   # one small D2H per row (a few hundred rows)
            capi.check(lib.oip_copy_d2h(ctx.h, C.c_void_p(got[k].ctypes.data), C.c_void_p(d_out_ptr + (int(g) - first) * row_b), row_b))
        ctx.sync()
        want = oracle.pan_rows(lambda i, a, b: synth.strip_dn(W, b - a, SEED + i, row0=a), N_CCD, W, kb_np, DX, DY, f, total_rows, chk)
        parity_bad = int((got != want).sum())
        parity_ok, parity_n = parity_bad == 0, int(chk.size)
        if not parity_ok:
            bad_rows = chk[np.flatnonzero((got != want).any(axis=1))]
            sys.stderr.write(f"rank {rank}: PARITY FAILURE: {parity_bad} px differ in rows {bad_rows[:20].tolist()}\n")

    # ---- e2e: host buffers through oip_pan_pipeline_host (H2D + kernels + D2H), every rank on its own block.  The host
    #      segments hold the block plus the halo rows (read from the "file" = regenerated, as a host reader would) and,
    #      where they lie outside, the stale rows of the partial last section
    e2e_ms, e2e_rows, e2e_in_bytes, e2e_note = 0.0, 0, 0, None
    if not args.no_e2e:
        import psutil
        need_full = N_CCD * W * 2 * rows + out_w * 2 * rows
        avail = psutil.virtual_memory().available // world
        e2e_rows = rows
        while e2e_rows > 4096 and (N_CCD * W * 2 + out_w * 2) * e2e_rows * 1.25 + (8 << 30) > avail:
            e2e_rows //= 2
        if e2e_rows != rows:
            e2e_note = f"host memory ({avail >> 30} GiB available per rank) holds {e2e_rows} of the block's {rows} lines"
        hd = ops.make_pan_desc([Shape(e2e_rows, W)] * N_CCD, ops.FMT_BE16, kb_host, DX, DY, [i > 0 for i in range(N_CCD)], f,
                               OutPtr(0, out_w), total_rows=total_rows, row0=first, n_rows=e2e_rows,
                               segs=[[(0, first, e2e_rows, W * 2)] for _ in range(N_CCD)])  # host segments filled in below
        pinned = []

        def pin(nbytes):
            p = C.c_void_p()
            capi.check(lib.oip_host_alloc_pinned(nbytes, C.byref(p)))
            pinned.append(p.value)
            return p.value

        chunk = 8192
        d_tmp = dev_alloc(chunk * W * 2)
        for i in range(N_CCD):
            (nf, nl), (sf, sl) = sharding.rows_needed(hd, i)
            segs = [(min(nf, first), max(nl, first + e2e_rows))]
            if sl > sf and not (segs[0][0] <= sf and sl <= segs[0][1]):
                segs.append((sf, sl))
            c = hd.ccd[i]
            c.n_seg = len(segs)
            for q, (a, b) in enumerate(segs):
                hp = pin((b - a) * W * 2)
                e2e_in_bytes += (b - a) * W * 2
                for r in range(a, b, chunk):   # the same synthetic rows, generated on the device and copied out
                    n = min(chunk, b - r)
                    capi.check(lib.oip_synth_strip_dn(ctx.h, d_tmp, W, n, r, W, SEED + i, 1))
                    capi.check(lib.oip_copy_d2h(ctx.h, C.c_void_p(hp + (r - a) * W * 2), C.c_void_p(d_tmp), n * W * 2))
                    ctx.sync()
                c.seg[q] = capi.RowSeg(hp, a, b - a, W * 2)
        capi.check(lib.oip_dev_free(ctx.h, C.c_void_p(d_tmp)))
        h_out = pin(e2e_rows * out_w * 2)
        hd.d_out = h_out

        def e2e_step():
            capi.check(lib.oip_pan_pipeline_host(ctx.h, C.byref(hd)))  # returns after the output landed in host memory

        e2e_step()
        sync_all()
        ts = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        e2e_ms = float(np.median(ts))  # the host link is shared with whatever else the box does: median, not mean
        # the host path and the device-resident path must agree: the host result goes back up in chunks and is compared there
        cmp_rows = 4096
        d_cmp = torch.empty((cmp_rows, out_w), dtype=torch.int16, device=dev)
        same = True
        for r in range(0, e2e_rows, cmp_rows):
            n = min(cmp_rows, e2e_rows - r)
            capi.check(lib.oip_copy_h2d(ctx.h, C.c_void_p(d_cmp.data_ptr()), C.c_void_p(h_out + r * out_w * 2), n * out_w * 2))
            ctx.sync()
            same = same and bool(torch.equal(d_cmp[:n], d_out[r:r + n].view(torch.int16)))
        for p_ in pinned:
            lib.oip_host_free_pinned(C.c_void_p(p_))
        if not same:
            parity_ok = False
            sys.stderr.write(f"rank {rank}: e2e output differs from the device-resident output\n")

    t = torch.tensor([total_ms, kernel_ms, e2e_ms, 0.0 if parity_ok else 1.0, float(clk["sm_mhz"] or 0)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([parity_n, parity_bad, e2e_rows, e2e_in_bytes], dtype=torch.int64, device=dev)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(cnt)
    total_ms, kernel_ms, e2e_max, parity_flag, _ = t.tolist()
    parity_n_all, parity_bad_all, e2e_rows_all, e2e_in_all = cnt.tolist()
    px_all = N_CCD * W * total_rows
    value = px_all * args.steps / (total_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    bpp = algorithmic_bytes_per_px()
    # the dominant kernel (pan_fast_kernel) covers fast_frac of the output; the generic kernel runs next to it on
    # a side stream, so the step time measured on the launching stream bounds the fast kernel's duration from above
    st = (C.c_int64 * 4)()
    capi.check(lib.oip_pan_plan_coverage(C.byref(desc), 1, 128, None, st))
    fast_frac = st[1] / max(1, st[0] + st[1])
    px_rank = N_CCD * W * rows
    achieved = px_rank * fast_frac * bpp / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "pan_kernel_traffic.json")) as fjs:
            tj = json.load(fjs)
            if world == 1 and tj.get("rows") == total_rows:
                traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass

    if rank == 0:
        clk["sm_mhz_min_over_ranks"] = tmin.tolist()[4]
        line = {
            "metric": METRIC, "value": value, "unit": "Gpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 RRC / f32 bicubic / u16", "data": "synthetic",
            "config": bench_config(world, total_rows),
            "gpu_launches": int(launches),
            "clocks": clk,
            "warmup_policy": f">= {WARM_SECONDS} s of GPU work before the timed steps (sustained clocks), {n_warm} steps",
            "parity_ok": bool(parity_flag == 0.0) if not args.no_parity or not args.no_e2e else None,
            "parity": {"rows_checked": int(parity_n_all), "px_checked": int(parity_n_all) * out_w, "px_different": int(parity_bad_all),
                       "against": "oracle.pan_rows (CPU restatement, pinned to the reference's compiled PreStitch); every rank: first/last "
                                  "16 rows of its block, +-8 rows around every section edge of both shifted CCDs, the stale bottom rows; "
                                  "plus e2e output == device-resident output"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "oip::panfast::pan_fast_kernel",
                         "algorithmic_bytes_per_px": bpp, "kernel_ms": kernel_ms, "px_share_of_step": fast_frac,
                         "note": "kernel_ms = CUDA-event time of one step on the launching stream (fast kernel + the "
                                 "generic-tile kernel joined from a side stream), max over ranks: an upper bound of the fast "
                                 "kernel's duration"},
        }
        if not args.no_e2e:
            px_e2e = N_CCD * W * e2e_rows_all
            line["e2e"] = {"value": px_e2e / (e2e_max * 1e-3) / 1e9, "unit": "Gpixel/s",
                           "h2d_bytes_per_step": int(e2e_in_all + world * N_CCD * W * 16),
                           "d2h_bytes_per_step": int(e2e_rows_all * out_w * 2), "ms_per_step": e2e_max,
                           "rows": int(e2e_rows_all), "note": e2e_note,
                           "api": "oip_pan_pipeline_host (C ABI, pinned host buffers), one call per rank on its own block"}
        else:
            line["e2e"] = {"value": None, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "--no-e2e"}
        if world == 1:
            # bounded CPU sample of the same workload (baseline, not the target)
            try:
                import cv2
                cpu_rows = 2048
                cpu_reference_pass(128)
                dt, _ = cpu_reference_pass(cpu_rows)
                line["cpu_baseline"] = {"value": N_CCD * W * cpu_rows / dt / 1e9, "unit": "Gpixel/s",
                                        "cores": cv2.getNumThreads(), "kind": cpu_kind(),
                                        "sample": f"{N_CCD}x{W}x{cpu_rows} lines of the same strip: RRC loop (1 thread) + "
                                                  f"cv2.remap cubic ({cv2.getNumThreads()} threads) + concat, in memory"}
            except Exception as e:  # the CPU leg must never take the GPU number down
                line["cpu_baseline"] = {"value": None, "unit": "Gpixel/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
            # second workload, same run: the C2 strip as raw DOWNLINK FILES (AOS frames -> IMTR frames -> image frames), every
            # stage-1 kernel inside the timed region (oip_downlink_to_stitched); its own roofline, parity against the oracle chain
            if not args.no_framed:
                try:
                    sys.path.insert(0, os.path.join(ROOT, "tools"))
                    import bench_framed
                    line["framed"] = bench_framed.run(ctx, 32768, max(3, min(args.steps, 10)), check=not args.no_parity, peak_gbs=peak)
                    if line["framed"].get("parity_ok") is False:
                        line["parity_ok"] = False
                except Exception as e:
                    line["framed"] = {"value": None, "note": f"failed: {e}"}
        emit(line)
    for p in opened:
        lib.oip_ipc_close(ctx.h, C.c_void_p(p))
    if world > 1:
        dist.barrier()
    for p in d_in:
        lib.oip_dev_free(ctx.h, C.c_void_p(p))
    if world > 1:
        dist.destroy_process_group()
    if not parity_flag == 0.0:
        raise SystemExit("parity check failed")


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """the ONE JSON line, written to the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    # anything a library prints on fd 1 (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=TOTAL_ROWS, help="strip length in lines (default: the C4 strip; smaller = smoke runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle spot check")
    ap.add_argument("--no-framed", action="store_true", help="skip the second workload (C2 as framed downlink files)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--rows", str(args.rows)] + \
                  (["--no-e2e"] if args.no_e2e else []) + (["--no-parity"] if args.no_parity else []) + (["--no-framed"] if args.no_framed else [])
            raise SystemExit(subprocess.call(cmd))
        run_gpu(args)


if __name__ == "__main__":
    main()
