"""torchrun worker: every rank computes its scanline-block shard of the fused PAN path, reading halo /
stale rows from the neighbours' HBM through CUDA-IPC mappings; rank 0 gathers and compares with the
CPU oracle of the WHOLE strip.  Launched by tests/test_multi_gpu.py (needs >= 2 GPUs)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opticalimageprocessor_b200 import capi, ops, sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    n, w, f, S, G = 3, 512, 20, 400, 450
    total = 1500 * world + 37
    dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
    first, last = sharding.shard_range(total, world, rank)
    rows = last - first
    full = [synth.strip_dn(w, total, 900 + i) for i in range(n)]       # every rank can regenerate any row
    kbs = [synth.rrc_coeffs(w, 950 + i) for i in range(n)]
    d_in, handles = [], []
    for i in range(n):
        p = C.c_void_p()
        nb = rows * w * 2
        capi.check(ctx.lib.oip_dev_alloc(ctx.h, nb, C.byref(p)))
        blk = np.ascontiguousarray(full[i][first:last].byteswap())
        capi.check(ctx.lib.oip_copy_h2d(ctx.h, p, C.c_void_p(blk.ctypes.data), nb))
        ctx.sync()
        d_in.append(p.value)
        hb = C.create_string_buffer(64)
        capi.check(ctx.lib.oip_ipc_export(ctx.h, p, hb))
        handles.append(hb.raw)
    allh = [None] * world
    dist.all_gather_object(allh, handles)
    d_kb = [torch.from_numpy(k).cuda() for k in kbs]
    out = torch.empty((rows, ops.pan_out_width(n, w, f)), dtype=torch.uint16, device="cuda")
    shape_only = [torch.empty((rows, w), dtype=torch.uint16) for _ in range(n)]
    desc = ops.make_pan_desc(shape_only, ops.FMT_BE16, d_kb, dX, dY, [i > 0 for i in range(n)], f, out,
                             total_rows=total, row0=first, n_rows=rows, section_rows=S, row_guard=G,
                             segs=[[(d_in[i], first, rows, w * 2)] for i in range(n)])
    opened = {}

    def peer(r, i):
        if (r, i) not in opened:
            p = C.c_void_p()
            capi.check(ctx.lib.oip_ipc_open(ctx.h, allh[r][i], C.byref(p)))
            opened[(r, i)] = p.value
        return opened[(r, i)]

    sharding.attach_segments(desc, n, total, world, rank, d_in, w * 2, peer)
    dist.barrier()
    capi.check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(desc)))
    capi.check(ctx.lib.oip_pan_check_error(ctx.h))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ok = True
    if True:
        import oracle
        want = oracle.pan_pipeline(full, kbs, dX, dY, f, S, G)[first:last]
        bad = np.argwhere(got != want)
        ok = bad.size == 0
        if not ok:
            print(f"rank {rank}: {len(bad)} px differ, first {bad[:5].tolist()}", flush=True)
    # ---- C3: the band alignment shards by section; every rank holds only the lines of its own sections and the strip's
    #      output is the concatenation of the ranks' outputs; the offset-estimation sums go through one all-reduce
    import oracle
    lines, wb, lps, ov = 2300 * world, 96, 700, 60
    mixed = np.random.default_rng(77).integers(0, 4096, (lines, 4 * wb)).astype(np.uint16)
    mkb = [synth.rrc_coeffs(wb, 300 + b) for b in range(4)]
    cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1)] for b in range(4)]
    cY = [[-3.2 + b, 2e-4 * (b + 1), -1e-8 * (b - 1.5)] for b in range(4)]
    secs = sharding.mss_sections(lines, lps, ov, 0, False, 200)
    mine = sharding.mss_rank_sections(secs, world, rank)
    lo, hi = mine[0][0], mine[-1][0] + mine[-1][1]
    shard = torch.from_numpy(np.ascontiguousarray(mixed[lo:hi])).cuda()
    n_out = sum(sc[4] for sc in mine)
    mout = torch.zeros((n_out, wb, 4), dtype=torch.uint16, device="cuda")
    k = ops.band_align_sections(ctx, shard, wb, [torch.from_numpy(q).cuda() for q in mkb], cX, cY, mine, secs, mout, total_lines=lines,
                                lines_per_section=lps, overlap=ov, min_process_lines=200, src_row0=lo)
    planes = [oracle.rrc(pl, q) for pl, q in zip(oracle.mss_split(mixed), mkb)]
    n_all, want_m = oracle.band_align(planes, cX, cY, lines_per_section=lps, overlap=ov, min_process_lines=200)
    o0 = mine[0][3]
    ok_m = k == n_out and np.array_equal(mout.cpu().numpy(), want_m[o0:o0 + n_out])
    if not ok_m:
        print(f"rank {rank}: MSS section shard differs from the whole-strip oracle", flush=True)
    ok = ok and ok_m
    # ---- N1 on shards (ref stitcher.h:148-201): oip_stt_parameters on every rank's own rows, the section that straddles
    #      the block boundary gathered onto one rank, ONE 4-double all-reduce; against the whole-strip CPU loop on cv2
    import cv2
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_phasecorr_cpu import _pair
    slines, sw, sov, sns, slps = 2400 * world, 512, 200, 2 * world - 1, 1000
    sa, sb = _pair(slines, sov, 1.37, -2.61, seed=4)
    pan1 = np.zeros((slines, sw), np.uint16); pan2 = np.zeros((slines, sw), np.uint16)
    pan1[:, sw - sov:] = sa
    pan2[:, :sov] = sb
    s0, s1 = sharding.shard_range(slines, world, rank)
    rows_g, mean_g = ops.calc_stt_parameters(ctx, torch.from_numpy(pan1[s0:s1].copy()).cuda(), torch.from_numpy(pan2[s0:s1].copy()).cuda(),
                                             overlap_cols=sov, edge_cols=6, sections=sns, lines_per_section=slps, total_lines=slines, row0=s0)

    def cvcorr(a, b):
        (x, y), r = cv2.phaseCorrelate(a, b)
        return x, y, r
    rows_w, mean_w = oracle.stt_parameters(pan1, pan2, overlap_cols=sov, edge_cols=6, sections=sns, lines_per_section=slps, correlate=cvcorr)
    owners = sharding.stt_section_owner(slines, sns, slps, world)
    ok_s = -1 in owners and mean_g is not None and all(abs(a - b) <= 2e-3 for a, b in zip(mean_g, mean_w))
    if not ok_s:
        print(f"rank {rank}: sharded offset estimate {mean_g} vs whole strip {mean_w} (owners {owners})", flush=True)
    ok = ok and ok_s
    # ---- stage 1 on byte-range shards: every rank scans its part of ONE downlink file (boundaries that cut
    #      through frames), the carries settle with one all-gather per round, the 3 counters go through an all-reduce,
    #      the IMTR cadence starts at the prefix of the all-gathered payload counts, the seq rules are combined on the host
    from test_sharding_cpu import _stage1_file
    fbuf = _stage1_file(prefix=b"\x00" * 13, restart_at=77)
    off_w, cnt_w = oracle.aos_scan(fbuf)
    imdt_w, st_w = oracle.imtr_deframe(fbuf, off_w)
    fa, fb = fbuf.size * rank // world + (3 if rank else 0), fbuf.size * (rank + 1) // world + (3 if rank + 1 < world else 0)
    fsub = torch.from_numpy(np.concatenate([fbuf[fa:min(fbuf.size, fb + 1023)], np.zeros(2 * 880, np.uint8)])).cuda()   # + room for 2 halo payloads
    n_sub = min(fbuf.size, fb + 1023) - fa

    def scan(carry):
        o, c, co = ops.aos_scan_shard(ctx, fsub[:n_sub], fb - fa, carry)
        return (o, c), co, int(c[0])
    (poff, pcnt), carry_in, n_valid_all = sharding.aos_resolve_carries(scan, world, rank)
    tc = torch.from_numpy(pcnt.copy()).cuda()
    dist.all_reduce(tc)
    heads = [None] * world
    mine_heads = [fsub[int(o):int(o) + 880].cpu().numpy() for o in poff[:2].tolist()]
    dist.all_gather_object(heads, mine_heads)
    f0, nf, skip, halo = sharding.imtr_shard_frames(n_valid_all, rank)
    following = [h for q in range(rank + 1, world) for h in heads[q]][:halo]
    offs = poff.tolist()
    for k, h in enumerate(following):
        fsub[n_sub + 880 * k:n_sub + 880 * (k + 1)] = torch.from_numpy(h).cuda()
        offs.append(n_sub + 880 * k)
    piece, info = ops.imtr_deframe_shard(ctx, fsub, torch.tensor(offs, dtype=torch.int64, device="cuda"), skip, nf)
    infos = [None] * world
    dist.all_gather_object(infos, info)
    keep, st_all = sharding.imtr_combine(infos)
    pieces = [None] * world
    dist.all_gather_object(pieces, piece.cpu().numpy() if keep[rank] else np.zeros(0, np.uint8))
    ok_f = tc.cpu().tolist() == cnt_w.tolist() and st_all == st_w.tolist() and np.array_equal(np.concatenate(pieces), imdt_w)
    gathered_off = [None] * world
    dist.all_gather_object(gathered_off, (poff.cpu().numpy().astype(np.uint64) + np.uint64(fa)).tolist())
    ok_f = ok_f and sum(gathered_off, []) == off_w.tolist() and (rank == 0 or carry_in > 0)
    if not ok_f:
        print(f"rank {rank}: sharded stage 1 differs from the whole-file oracle (counters {tc.cpu().tolist()} vs {cnt_w.tolist()}, stats {st_all} vs {st_w.tolist()})", flush=True)
    ok = ok and ok_f
    # ---- frame index over the IMDT pieces, which stay on the GPUs that produced them: every rank searches its piece + the
    #      174-byte head of what follows, one all_gather_object of the (offset, trailer) pairs, the host chain on every rank
    my_piece = piece if keep[rank] else piece[:0]
    sizes = [int(q["imdt_bytes"]) if keep[r] else 0 for r, q in enumerate(infos)]
    hd = [None] * world
    dist.all_gather_object(hd, my_piece[:sharding.FRAME_HALO].cpu().numpy().tobytes())
    halo_b = sharding.frames_piece_halo([np.frombuffer(h, np.uint8) for h in hd], rank)
    ext = torch.cat([my_piece, torch.from_numpy(halo_b.copy()).cuda()]) if halo_b.size else my_piece
    ents_g, fst_g = sharding.frames_index_shards(lambda: ops.image_frames_hits(ctx, ext), sizes, rank, 32, 8)
    ents_1, fst_1 = ops.image_frames_index(ctx, torch.from_numpy(imdt_w).cuda(), 32, 8)
    fst_w = oracle.image_frames(imdt_w, 32, 8)[4]
    ok_i = sum(sizes) == imdt_w.size and fst_g.tolist() == fst_w.tolist() == fst_1.tolist() and int(fst_g[1]) > 0 and all(
        ents_g[f].frame_off == ents_1[f].frame_off and list(ents_g[f].tile_off) == list(ents_1[f].tile_off) for f in range(int(fst_g[1])))
    if not ok_i:
        print(f"rank {rank}: frame index over pieces {fst_g.tolist()} vs whole stream {fst_w.tolist()} (sizes {sizes})", flush=True)
    ok = ok and ok_i
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    for p in opened.values():
        ctx.lib.oip_ipc_close(ctx.h, C.c_void_p(p))
    dist.barrier()
    for p in d_in:
        ctx.lib.oip_dev_free(ctx.h, C.c_void_p(p))
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if t.item() == 1.0 else "FAIL", f"world={world} peers_opened_rank0={sorted(opened)}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
