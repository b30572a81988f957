#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused PAN hot path (unpack -> RRC -> shift -> stitch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the fused kernel over one strip shard: BASELINE.json configs[1]
("3-CCD panchromatic strip, 3x8192 px with overlap, 32k lines") per GPU.  At N > 1 the strip is
N x 32768 lines long, sharded by scanline blocks (weak scaling); section geometry is global and
the few halo rows a shard needs from its neighbours are read by the kernel straight from the
neighbour GPU's memory over NVLink (CUDA IPC peer mappings), no data-path collective.

value    : Gpixel/s, device-timed (CUDA events on the launching stream), inputs resident in HBM,
           whole job over all ranks, max over ranks.
e2e      : same metric through the host-buffer C-ABI call (pinned host in/out, H2D/D2H inside).
roofline : algorithmic HBM bytes of the fused kernel / its mean launch time vs the measured copy peak.
cpu_baseline / --impl reference : the reference's CPU path for the same work = fp64 RRC loop
           (single thread, ref imageop.h:129-138) + cv2.remap in 30000-row sections with OpenCV's own
           thread pool (the library call at imageop.h:258) + line concat, on a bounded slice.
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

N_CCD, W, ROWS, FOLD = 3, 8192, 32768, 200
DX = [0.0, 1.37, -0.83]
DY = [0.0, -2.61, 3.19]
SEED = 0x0A11CE02
METRIC = "Gpixel/s (raw->stitched, device-timed)"
WORKLOAD = "C2: 3 CCD x 8192 px x 32768 lines per GPU, 16-bit BE raw in, fold 200, RRC + cubic shift + stitch"


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
    """the reference's CPU path on `rows` lines of the C2 strip; returns seconds (in-memory, no disk)"""
    import cv2
    import oracle
    from opticalimageprocessor_b200 import synth
    oracle.build()
    if threads:
        cv2.setNumThreads(threads)
    S = 30000
    ccds = [synth.strip_dn(W, rows, SEED + i) for i in range(N_CCD)]
    kbs = [synth.rrc_coeffs(W, SEED + 100 + i) for i in range(N_CCD)]
    mx = [(np.arange(W)[None, :] + np.zeros((min(S, rows), 1)) + DX[i]).astype(np.float32) for i in range(N_CCD)]
    my = [(np.arange(min(S, rows))[:, None] + np.zeros((1, W)) + DY[i]).astype(np.float32) for i in range(N_CCD)]
    f = FOLD // 2
    rrc = oracle.ref_inplace_rrc if oracle.ref_oip_lib() is not None else oracle.rrc   # the reference's own compiled loop if built
    t0 = time.perf_counter()
    parts = []
    for i in range(N_CCD):
        r = rrc(ccds[i], kbs[i])                                          # ref imageop.h:129-138 (1 thread)
        if i > 0:                                                         # ref stitcher.h:83-139 / imageop.h:258
            r = cv2.remap(r, mx[i][:rows], my[i][:rows], cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
        lo, hi = (0 if i == 0 else f), (W if i == N_CCD - 1 else W - f)
        parts.append(r[:, lo:hi])
    out = np.concatenate(parts, axis=1)                                   # ref imageop.h:340-355
    dt = time.perf_counter() - t0
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
    import cv2
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 4096
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_pass(256)
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
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 RRC / f32 bicubic / u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpu_sample": f"{rows} lines of the same strip per step"},
        "cpu_baseline": {"value": val, "unit": "Gpixel/s", "cores": cores, "kind": cpu_kind(),
                         "sample": f"{N_CCD}x{W}x{rows} lines: IMO::InplaceRRC (the reference's imageop.h compiled under "
                                   f"oracle/_ref when present, else its C restatement; 1 thread like the reference) + cv2.remap "
                                   f"INTER_CUBIC ({cores} OpenCV threads, the reference's own library call) + concat, in memory"},
        "e2e": {"value": val, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from opticalimageprocessor_b200 import build, capi, ops, synth
    import ctypes as C

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
    dev = torch.device("cuda", local)
    total_rows = ROWS * world
    row0 = ROWS * rank
    out_w = ops.pan_out_width(N_CCD, W, FOLD // 2)

    # ---- synthetic shard (host, pinned) -> device.  BE16 = byte order inside the downlink sub-images
    host_in = []
    for i in range(N_CCD):
        a = synth.strip_dn(W, ROWS, SEED + i, row0=row0).byteswap()
        t = torch.from_numpy(a).pin_memory()
        host_in.append(t)
    kb_np = [synth.rrc_coeffs(W, SEED + 100 + i) for i in range(N_CCD)]
    kb_host = [torch.from_numpy(k) for k in kb_np]
    # raw cudaMalloc allocations (not torch's caching allocator): an IPC handle maps a whole allocation
    d_in = []
    for t in host_in:
        p = C.c_void_p()
        capi.check(ctx.lib.oip_dev_alloc(ctx.h, t.numel() * 2, C.byref(p)))
        capi.check(ctx.lib.oip_copy_h2d(ctx.h, p, C.c_void_p(t.data_ptr()), t.numel() * 2))
        d_in.append(p.value)
    ctx.sync()
    d_kb = [t.to(dev) for t in kb_host]
    d_out = torch.empty((ROWS, out_w), dtype=torch.uint16, device=dev)
    host_out = torch.empty((ROWS, out_w), dtype=torch.uint16).pin_memory()

    # ---- halo rows from the neighbour shards: peer mappings over NVLink (no collective on the data path)
    desc = ops.make_pan_desc(host_in, ops.FMT_BE16, d_kb, DX, DY, [i > 0 for i in range(N_CCD)], FOLD // 2, d_out,
                             total_rows=total_rows, row0=row0, n_rows=ROWS,
                             segs=[[(d_in[i], row0, ROWS, W * 2)] for i in range(N_CCD)])
    opened = []
    if world > 1:
        handles = []
        for i in range(N_CCD):
            hb = C.create_string_buffer(64)
            capi.check(ctx.lib.oip_ipc_export(ctx.h, C.c_void_p(d_in[i]), hb))
            handles.append(hb.raw)
        allh = [None] * world
        dist.all_gather_object(allh, handles)
        peer_ptr = {}

        def peer(r, i):
            if (r, i) not in peer_ptr:
                p = C.c_void_p()
                capi.check(ctx.lib.oip_ipc_open(ctx.h, allh[r][i], C.byref(p)))
                peer_ptr[(r, i)] = p.value
                opened.append(p.value)
            return peer_ptr[(r, i)]

        from opticalimageprocessor_b200 import sharding
        sharding.attach_segments(desc, N_CCD, total_rows, world, rank, d_in, W * 2, peer)
        dist.barrier()

    def step():
        capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(desc)))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    capi.check(ctx.lib.oip_pan_check_error(ctx.h))
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
    capi.check(ctx.lib.oip_pan_check_error(ctx.h))

    # ---- e2e: host buffers through oip_pan_pipeline_host (H2D + kernels + D2H), own shard only
    def e2e_step():
        ops.pan_pipeline_host(ctx, host_in, kb_host, DX, DY, FOLD // 2, host_out, fmt=ops.FMT_BE16)

    e2e_ms = None
    if world == 1:
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 7))
        ts = []
        for _ in range(n_e2e):
            t0 = time.perf_counter()
            e2e_step()  # returns after the output block landed in host memory
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        e2e_ms = float(np.median(ts))  # the host link is shared with whatever else the box does: median, not mean
        # the device-resident path and the host path must agree
        if not torch.equal(host_out.view(torch.int16), d_out.cpu().view(torch.int16)):
            raise SystemExit("e2e output differs from the device-resident output")

    t = torch.tensor([total_ms, kernel_ms, e2e_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms, e2e_max = t.tolist()
    px_rank = N_CCD * W * ROWS
    px_all = px_rank * world
    value = px_all * args.steps / (total_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    bpp = algorithmic_bytes_per_px()
    # the dominant kernel (pan_fast_kernel) covers fast_frac of the output; the generic kernel runs next to it on
    # a side stream, so the step time measured on the launching stream bounds the fast kernel's duration from above
    st = (C.c_int64 * 4)()
    capi.check(ctx.lib.oip_pan_plan_coverage(C.byref(desc), 1, 128, None, st))
    fast_frac = st[1] / max(1, st[0] + st[1])
    achieved = px_rank * fast_frac * bpp / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "pan_kernel_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("workload") == WORKLOAD:
                traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 RRC / f32 bicubic / u16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_ccd": N_CCD, "w": W, "rows_per_gpu": ROWS, "total_rows": total_rows,
                       "fold_cols": FOLD, "dX": DX, "dY": DY, "l2": "inputs (1.6 GB) and output (1.6 GB) per step >> 126 MB L2",
                       "multi_gpu": "scanline-block shards, halo rows read from peer HBM over NVLink (CUDA IPC)" if world > 1 else "single GPU"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "oip::panfast::pan_fast_kernel",
                         "algorithmic_bytes_per_px": bpp, "kernel_ms": kernel_ms, "px_share_of_step": fast_frac,
                         "note": "kernel_ms = CUDA-event time of one step on the launching stream (fast kernel + the "
                                 "generic-tile kernel joined from a side stream): an upper bound of the fast kernel's duration"},
        }
        if world == 1:
            line["e2e"] = {"value": px_rank / (e2e_max * 1e-3) / 1e9, "unit": "Gpixel/s",
                           "h2d_bytes_per_step": int(sum(t.numel() * 2 for t in host_in) + sum(k.numel() * 8 for k in kb_host)),
                           "d2h_bytes_per_step": int(host_out.numel() * 2), "ms_per_step": e2e_max,
                           "api": "oip_pan_pipeline_host (C ABI, pinned host buffers)"}
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
        else:
            line["e2e"] = {"value": None, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "measured at N=1 only"}
        emit(line)
    for p in opened:
        ctx.lib.oip_ipc_close(ctx.h, C.c_void_p(p))
    if world > 1:
        dist.barrier()
    for p in d_in:
        ctx.lib.oip_dev_free(ctx.h, C.c_void_p(p))
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup)]
            raise SystemExit(subprocess.call(cmd))
        run_gpu(args)


if __name__ == "__main__":
    main()
