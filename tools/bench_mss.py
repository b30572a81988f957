"""C3 (BASELINE configs[2]): 4-band MSS band alignment, 65536 lines x 12288 px (4 x 3072) per GPU, device-timed; under
torchrun every rank aligns the sections of its own strip shard (the path shards by section: no halo, no collective) and
rank 0 prints one JSON line with the whole-job throughput (max over ranks).
    python tools/bench_mss.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_mss.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from opticalimageprocessor_b200 import ops, sharding, synth

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ops.Context(local)
if os.environ.get("MSS_FAST_ROWS"):  # development: warp-tile height of mss_fast_kernel
    ctx.set_option("mss_fast_rows", int(os.environ["MSS_FAST_ROWS"]))
LINES, WB, LPS, OV = 65536, 3072, 20000, 520
total_lines = LINES * world                                   # weak scaling: one C3-sized strip per GPU
secs = sharding.mss_sections(total_lines, LPS, OV, 0, False, 1500)
mine = sharding.mss_rank_sections(secs, world, rank)
lo, hi = mine[0][0], mine[-1][0] + mine[-1][1]
g = torch.Generator(device="cuda").manual_seed(100 + rank)
mss = torch.empty((hi - lo, 4 * WB), dtype=torch.uint16, device="cuda")
for r0 in range(0, hi - lo, 16384):
    r1 = min(hi - lo, r0 + 16384)
    mss[r0:r1] = torch.randint(64, 4032, (r1 - r0, 4 * WB), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16)
kbs = [torch.from_numpy(synth.rrc_coeffs(WB, 60 + b)).cuda() for b in range(4)]
cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
cY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8 * (b - 1.5)] for b in range(4)]
n_out = sum(s[4] for s in mine)
out = torch.zeros((n_out, WB, 4), dtype=torch.uint16, device="cuda")


def step():
    ops.band_align_sections(ctx, mss, WB, kbs, cX, cY, mine, secs, out, total_lines=total_lines, lines_per_section=LPS, overlap=OV,
                            src_row0=lo)


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
px = sum(s[1] for s in secs) * 4 * WB                       # source samples the sections read (overlap lines twice, as in the reference)
if rank == 0:
    print(json.dumps({"workload": "C3: 4-band MSS band alignment, 65536 lines x 4 x 3072 px per GPU, sections of 20000 lines / overlap 520",
                      "n_gpus": world, "ms_per_step": ms, "Gpixel_per_s": px / ms / 1e6, "sections": len(secs),
                      "sections_per_rank": [len(sharding.mss_rank_sections(secs, world, r)) for r in range(world)],
                      "scaling": "weak", "collectives": 0}), flush=True)
if world > 1:
    dist.destroy_process_group()
