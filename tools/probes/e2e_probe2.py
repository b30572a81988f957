"""development aid: the copy pattern of oip_pan_pipeline_host done with torch copies on the SAME pinned buffers bench.py uses
(numpy-born) and on fresh torch.empty pinned buffers -- separates "pipeline structure" from "where the host pages live"."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from opticalimageprocessor_b200 import ops, synth
W, ROWS = bench.W, bench.ROWS
out_w = ops.pan_out_width(3, W, bench.FOLD // 2)
def run(host_in, host_out, blk_rows, tag):
    s1, s2, sc = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    din = [[torch.empty((blk_rows, W), dtype=torch.uint16, device="cuda") for _ in range(3)] for _ in range(3)]
    dout = [torch.empty((blk_rows, out_w), dtype=torch.uint16, device="cuda") for _ in range(3)]
    def once():
        for b in range(ROWS // blk_rows):
            s = b % 3
            r0, r1 = b * blk_rows, (b + 1) * blk_rows
            with torch.cuda.stream(s1):
                for i in range(3): din[s][i].copy_(host_in[i][r0:r1], non_blocking=True)
                e_in = torch.cuda.Event(); e_in.record(s1)
            with torch.cuda.stream(sc):
                sc.wait_event(e_in)
                e_c = torch.cuda.Event(); e_c.record(sc)
            with torch.cuda.stream(s2):
                s2.wait_event(e_c)
                host_out[r0:r1].copy_(dout[s], non_blocking=True)
        torch.cuda.synchronize()
    once()
    t0 = time.perf_counter()
    for _ in range(3): once()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"{tag}, {blk_rows}-row blocks: {ms:.1f} ms  ({3*W*ROWS*2/ms/1e6:.1f} GB/s up, {ROWS*out_w*2/ms/1e6:.1f} GB/s down)", flush=True)
a_in = [torch.from_numpy(synth.strip_dn(W, ROWS, bench.SEED + i).byteswap()).pin_memory() for i in range(3)]
a_out = torch.empty((ROWS, out_w), dtype=torch.uint16).pin_memory()
b_in = [torch.empty((ROWS, W), dtype=torch.uint16).pin_memory() for _ in range(3)]
b_out = torch.empty((ROWS, out_w), dtype=torch.uint16).pin_memory()
for blk in (2048, 4096):
    run(a_in, a_out, blk, "numpy-born pinned buffers")
    run(b_in, b_out, blk, "torch.empty pinned buffers")
import subprocess
print(subprocess.run("numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -6; nproc", shell=True, capture_output=True, text=True).stdout)
