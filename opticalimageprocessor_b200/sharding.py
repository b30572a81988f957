"""Scanline-block sharding of a strip across the GPUs of one box.

No data-path collective: a rank owns `rows_per_rank` consecutive lines of every CCD; the few rows
its shifted tiles read beyond that block (halo rows, and for the last rank the stale rows of the
previous 30000-row section) are supplied as extra row segments that point into the owning
neighbour's buffer.  On a GPU box those pointers are CUDA-IPC mappings (read over NVLink by the
kernel's bulk copies); `torch.distributed` only carries the handle exchange.

The planning here is pure host logic (it only calls the library's `oip_pan_rows_needed`, which needs
no GPU), so it is covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Sequence, Tuple

from . import capi


def shard_range(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, last) lines owned by `rank`: equal blocks, the last rank takes the remainder"""
    per = total_rows // world
    first = rank * per
    last = total_rows if rank == world - 1 else first + per
    return first, last


def rows_needed(desc: capi.PanDesc, ccd: int) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    L = capi.load()
    f, l, sf, sl = (C.c_int64() for _ in range(4))
    capi.check(L.oip_pan_rows_needed(C.byref(desc), ccd, f, l, sf, sl))
    return (f.value, l.value), (sf.value, sl.value)


def peer_requirements(desc: capi.PanDesc, n_ccd: int, total_rows: int, world: int, rank: int) -> List[List[int]]:
    """for each CCD, the ranks (other than `rank`) that own rows this shard reads"""
    out = []
    for i in range(n_ccd):
        need = rows_needed(desc, i)
        owners = []
        for a, b in need:
            if b <= a:
                continue
            for r in range(world):
                if r == rank:
                    continue
                lo, hi = shard_range(total_rows, world, r)
                if min(b, hi) > max(a, lo) and r not in owners:
                    owners.append(r)
        out.append(owners)
    return out


def attach_segments(desc: capi.PanDesc, n_ccd: int, total_rows: int, world: int, rank: int, own_ptr: Sequence[int],
                    pitch_bytes: int, peer_ptr: Callable[[int, int], int]) -> Dict[int, List[int]]:
    """fill desc.ccd[i].seg[] with the own block plus one segment per neighbour that owns needed rows.
    peer_ptr(rank, ccd) -> device-visible pointer to that rank's block of CCD `ccd`."""
    req = peer_requirements(desc, n_ccd, total_rows, world, rank)
    for i in range(n_ccd):
        first, last = shard_range(total_rows, world, rank)
        segs = [(own_ptr[i], first, last - first)]
        for r in req[i]:
            lo, hi = shard_range(total_rows, world, r)
            segs.append((peer_ptr(r, i), lo, hi - lo))
        if len(segs) > capi.MAX_SEG:
            raise ValueError(f"CCD {i}: shard needs rows from {len(segs) - 1} neighbours (max {capi.MAX_SEG - 1})")
        c = desc.ccd[i]
        c.n_seg = len(segs)
        for s, (base, r0, nr) in enumerate(segs):
            c.seg[s] = capi.RowSeg(base, r0, nr, pitch_bytes)
    return {i: req[i] for i in range(n_ccd)}


def covers(desc: capi.PanDesc, ccd: int) -> bool:
    """every row the shard reads is inside one of the attached segments"""
    c = desc.ccd[ccd]
    for a, b in rows_needed(desc, ccd):
        for g in range(a, b):
            if not any(c.seg[s].row0 <= g < c.seg[s].row0 + c.seg[s].n_rows for s in range(c.n_seg)):
                return False
    return True


# ------------------------------------------------------------------------------------------------
# N1 (ref stitcher.h:148-201): the sections of the inter-CMOS offset estimate on a sharded strip
# ------------------------------------------------------------------------------------------------
def stt_section_offsets(total_lines: int, sections: int, lines_per_section: int) -> List[int]:
    """first line of every section (ref stitcher.h:151-152, :167)"""
    gap = (total_lines - sections * lines_per_section) // (sections + 1)
    return [gap + i * (gap + lines_per_section) for i in range(sections)]


def stt_section_owner(total_lines: int, sections: int, lines_per_section: int, world: int) -> List[int]:
    """rank whose scanline block holds a section entirely, or -1 when it straddles two blocks (oip_stt_parameters marks
    such a section valid = -1; stt_gather_straddling hands its rows to one rank)"""
    owners = []
    for off in stt_section_offsets(total_lines, sections, lines_per_section):
        own = -1
        for r in range(world):
            lo, hi = shard_range(total_lines, world, r)
            if lo <= off and off + lines_per_section <= hi:
                own = r
        owners.append(own)
    return owners


def stt_section_plan(total_lines: int, sections: int, lines_per_section: int, ranges: Sequence[Tuple[int, int]]):
    """for every section of the offset estimate: (first line, owner rank, [(rank, first, last) pieces in row order]).
    ranges[r] = [first, last) lines held by rank r.  A section inside one block has one piece; a section that straddles
    blocks is owned by the rank that holds its first line and the other ranks send their rows of the two overlap
    slices (lines x (overlap - edge) px x 2 strips: a few MB) -- ref stitcher.h:166-199 correlates every section."""
    plan = []
    for off in stt_section_offsets(total_lines, sections, lines_per_section):
        pieces = []
        for r, (lo, hi) in enumerate(ranges):
            a, b = max(lo, off), min(hi, off + lines_per_section)
            if b > a:
                pieces.append((r, a, b))
        pieces.sort(key=lambda p: p[1])
        covered = sum(b - a for _, a, b in pieces)
        if covered != lines_per_section or not pieces:
            raise ValueError(f"section at line {off}: the shards hold {covered} of its {lines_per_section} lines")
        plan.append((off, pieces[0][0], pieces))
    return plan


def stt_gather_straddling(pan1, pan2, row0: int, rank: int, plan, cols1: Tuple[int, int], cols2: Tuple[int, int], group=None):
    """exchange step of the sharded offset estimate: for every section that straddles scanline blocks, the non-owner
    ranks send their rows of pan1[:, cols1] / pan2[:, cols2] to the owner (point-to-point, torch.distributed: NCCL over
    NVLink on a GPU box, gloo in the CPU tests).  Returns [(section index, a, b)] on the owner: the assembled
    lines_per_section x cols slices, ready for the phase correlation."""
    import torch
    import torch.distributed as dist
    out = []
    n1, n2 = cols1[1] - cols1[0], cols2[1] - cols2[0]
    assert n1 == n2
    for idx, (off, owner, pieces) in enumerate(plan):
        if len(pieces) == 1:
            continue
        mine = [p for p in pieces if p[0] == rank]
        if rank == owner:
            lps = pieces[-1][2] - off
            buf = torch.empty((2, lps, n1), dtype=torch.int16, device=pan1.device)
            for r, a, b in pieces:
                dst = buf[:, a - off:b - off]
                if r == rank:
                    dst[0].copy_(pan1[a - row0:b - row0, cols1[0]:cols1[1]].view(torch.int16))
                    dst[1].copy_(pan2[a - row0:b - row0, cols2[0]:cols2[1]].view(torch.int16))
                else:
                    tmp = torch.empty((2, b - a, n1), dtype=torch.int16, device=pan1.device)
                    # bytes on the wire: NCCL has no 16-bit integer type
                    dist.recv(tmp.view(torch.uint8), src=r if group is None else dist.get_global_rank(group, r), group=group)
                    dst.copy_(tmp)
            out.append((idx, buf[0].view(torch.uint16), buf[1].view(torch.uint16)))
        elif mine:
            _, a, b = mine[0]
            tmp = torch.stack([pan1[a - row0:b - row0, cols1[0]:cols1[1]].view(torch.int16),
                               pan2[a - row0:b - row0, cols2[0]:cols2[1]].view(torch.int16)]).contiguous()
            dist.send(tmp.view(torch.uint8), dst=owner if group is None else dist.get_global_rank(group, owner), group=group)
    return out


def stt_combine(sums: Sequence[float]):
    """(dx, dy, response) means from the all-reduced {sum dx, sum dy, sum response, n valid} (ref stitcher.h:197-199);
    None when no section was valid (the reference throws)"""
    return None if sums[3] == 0 else (sums[0] / sums[3], sums[1] / sums[3], sums[2] / sums[3])


# ------------------------------------------------------------------------------------------------
# Band alignment (ref preproc.h:351-408): its sections are independent -- each one is remapped from its own source
# lines with section-local map coordinates -- so the MSS path shards by SECTION: no halo rows, no exchange.
# ------------------------------------------------------------------------------------------------
def mss_sections(lines: int, lines_per_section: int = 20000, overlap: int = 520, line_offset: int = 0,
                 keep_leading: bool = False, min_process_lines: int = 1500):
    """the reference's section loop (ref preproc.h:379-408) as (src_row0, n_src_rows, rows_dropped, out_row0, n_out_rows)"""
    secs, offset, processed, i = [], line_offset, 0, 0
    while True:
        n = min(lines - offset, lines_per_section)                                  # :380
        if lines < offset or n < min_process_lines:                                  # :381
            break
        y0 = 0 if (i == 0 and keep_leading) else overlap                             # :392-402
        secs.append((offset, n, y0, processed, n - y0))
        processed += n - y0                                                          # :396, :405
        offset += lines_per_section - overlap                                        # :407
        i += 1
    return secs


def mss_rank_sections(secs, world: int, rank: int):
    """contiguous runs of sections, balanced by output rows"""
    total = sum(s[4] for s in secs)
    out, acc = [], 0
    for s in secs:
        mid = acc + s[4] / 2.0
        if int(mid * world / max(total, 1)) == rank:
            out.append(s)
        acc += s[4]
    return out


# ------------------------------------------------------------------------------------------------
# Stage 1 on byte-range shards of one downlink file (ref aux_separator.h:395-467 is one sequential scan)
# ------------------------------------------------------------------------------------------------
AOS_FRAME = 1024


def aos_shard_ranges(n_bytes: int, world: int, phase: int = 0):
    """[first, last) file bytes whose sync candidates rank r owns: equal runs of whole 1024-byte slots at the cadence phase
    of the file (the offset of its first sync word), so that in a clean downlink no frame straddles a boundary and every
    shard's scan starts at its first byte; rank 0 also owns the bytes in front of the phase, the last rank the tail.
    A rank READS [first, min(n_bytes, last + 1023)): the halo lets it validate the frames that start near its end."""
    phase = phase % AOS_FRAME if n_bytes > phase else 0
    slots = max(0, (n_bytes - phase) // AOS_FRAME)
    out = []
    for r in range(world):
        a = 0 if r == 0 else phase + (slots * r // world) * AOS_FRAME
        b = n_bytes if r == world - 1 else phase + (slots * (r + 1) // world) * AOS_FRAME
        out.append((a, max(a, b)))
    return out


def aos_resolve_carries(scan_shard, world: int, rank: int, group=None, max_rounds: int = None):
    """Runs the shard scans with ONE all-gather per round.  scan_shard(carry_in) -> (result, carry_out, n_valid)
    scans this rank's shard from buffer offset carry_in.  Round 0 assumes carry_in = 0 everywhere (true whenever the
    boundaries sit on the frame cadence); the all-gather of (carry_in used, carry_out, n_valid) shows every rank the carry
    its left neighbour produced; a rank whose assumption was wrong scans again.  A scan re-synchronises at the first frame
    that no earlier frame overlaps, so carry_out practically never depends on carry_in and the second round is final;
    the loop is bounded by `world` rounds for the pathological chain.
    Returns (result, carry_in, [n_valid of every rank])."""
    import torch
    import torch.distributed as dist
    carry_in = 0
    res, carry_out, n_valid = scan_shard(carry_in)
    rounds = world if max_rounds is None else max_rounds
    for _ in range(rounds):
        mine = torch.tensor([carry_in, carry_out, n_valid], dtype=torch.int64)
        allv = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        if world > 1:
            dev = None
            if dist.get_backend(group) == "nccl":
                dev = torch.device("cuda", torch.cuda.current_device())
                mine = mine.to(dev)
                allv = [t.to(dev) for t in allv]
            dist.all_gather(allv, mine, group=group)
            allv = [t.cpu() for t in allv]
        else:
            allv = [mine]
        want = 0 if rank == 0 else int(allv[rank - 1][1])
        consistent = all((0 if r == 0 else int(allv[r - 1][1])) == int(allv[r][0]) for r in range(world))
        if consistent:
            return res, carry_in, [int(t[2]) for t in allv]
        if want != carry_in:
            carry_in = want
            res, carry_out, n_valid = scan_shard(carry_in)
    raise RuntimeError("AOS shard carries did not settle")


def imtr_shard_frames(n_valid_all, rank: int):
    """The fixed 882-byte cadence over the concatenated 880-byte payloads (ref aux_separator.h:487-510) on shards: rank r's
    payloads are stream bytes [880 * P_r, 880 * P_{r+1}) with P_r = payloads of the ranks before it (prefix of the
    all-gathered n_valid).  It owns the frames that START in its range: returns (first_frame, n_frames, skip, halo_payloads):
    skip = bytes of its first payload that still belong to the previous rank's last frame, halo_payloads = payloads of the
    following ranks its last frame reaches into (0..2)."""
    total = sum(n_valid_all)
    p0 = sum(n_valid_all[:rank])
    p1 = p0 + n_valid_all[rank]
    n_frames_global = total * 880 // 882
    f0 = (880 * p0 + 881) // 882
    f1 = min((880 * p1 + 881) // 882, n_frames_global)
    n = max(0, f1 - f0)
    skip = 882 * f0 - 880 * p0 if n else 0
    halo = 0
    if n:
        end = 882 * f1                       # one past the last stream byte of my last frame
        halo = max(0, (end - 880 * p1 + 879) // 880)
    return f0, n, skip, halo


def imtr_combine(infos):
    """Puts the shards of the IMTR re-framing together (ref aux_separator.h:513-533: the sequence rules look at the
    previously ACCEPTED frame, which for a shard's first valid frame lives on another rank).
    infos[r] = dict(n_frames, n_valid, bad=[sig, endsig, type, crc], first_seq, last_seq, gaps, restarts (both WITHOUT the
    rule of the shard's first valid frame), local_restart (index of the last restart among its valid frames, -1 none),
    first_chid, imdt_bytes) -- what every rank all-gathers (a dozen integers).
    Returns (keep[r]: does rank r's IMDT piece belong to the product, stats[9] of the whole stream)."""
    prev, gaps, restarts, last_restart_rank = 0, 0, 0, -1
    for r, q in enumerate(infos):
        if q["n_valid"] == 0:
            continue
        boundary_restart = prev == 0                                   # :513-528, lastImtrSeq == 0
        restarts += int(boundary_restart) + q["restarts"]
        gaps += int(prev + 1 != q["first_seq"]) + q["gaps"]            # :530-533
        if boundary_restart or q["local_restart"] >= 0:
            last_restart_rank = r
        prev = q["last_seq"]
    keep = [q["n_valid"] > 0 and r >= last_restart_rank for r, q in enumerate(infos)]
    n_valid = sum(q["n_valid"] for q in infos)
    stats = [sum(q["n_frames"] for q in infos), n_valid] + [sum(q["bad"][k] for q in infos) for k in range(4)] + \
            [gaps, infos[last_restart_rank]["first_chid"] if last_restart_rank >= 0 else -1, restarts]
    return keep, stats


# ------------------------------------------------------------------------------------------------
# Frame index over the pieces of an IMDT stream (the output of the sharded IMTR re-framing stays where it was produced)
# ------------------------------------------------------------------------------------------------
FRAME_TRAILER = 172
FRAME_HALO = FRAME_TRAILER + 2    # bytes of the following pieces a rank reads: a signature that starts in its last 3 bytes
                                  # still ends, with its trailer, inside piece + halo


def frames_piece_halo(heads: Sequence, rank: int):
    """the first FRAME_HALO bytes of the stream that follows rank's piece, cut from the all-gathered heads of the pieces
    (heads[r] = the first FRAME_HALO bytes of piece r, shorter for a short piece; a piece shorter than the halo lets the
    next one contribute)"""
    import numpy as np
    out = []
    need = FRAME_HALO
    for q in range(rank + 1, len(heads)):
        if need <= 0:
            break
        h = np.asarray(heads[q], np.uint8)[:need]
        out.append(h)
        need -= h.size
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


def frames_local_hits(find_hits, piece_bytes: Sequence[int], rank: int):
    """what a rank contributes to the exchange: the signatures that START inside its piece, as (global offsets, trailer bytes).
    find_hits() -> (offsets, trailers) of the signatures in this rank's piece + halo, offsets relative to the piece."""
    import numpy as np
    off, tr = find_hits()
    off = np.asarray(off, np.uint64)
    tr = np.asarray(tr, np.uint8).reshape(-1, FRAME_TRAILER)
    mine = off < np.uint64(piece_bytes[rank])
    base = sum(int(b) for b in piece_bytes[:rank])
    return (off[mine] + np.uint64(base)).tolist(), tr[mine].tobytes()


def frames_chain_all(payloads, piece_bytes: Sequence[int], tile_cols: int, tile_lines: int):
    """the host chain over the all-gathered contributions (rank order = stream order) -> (entries, stats[4])"""
    import numpy as np
    from . import ops
    hits = np.array(sum((list(p[0]) for p in payloads), []), np.uint64)
    trailers = np.frombuffer(b"".join(p[1] for p in payloads), np.uint8).reshape(-1, FRAME_TRAILER)
    return ops.image_frames_chain(hits, trailers, sum(int(b) for b in piece_bytes), tile_cols, tile_lines)


def frames_index_shards(find_hits, piece_bytes: Sequence[int], rank: int, tile_cols: int, tile_lines: int, group=None):
    """oip_image_frames_index for an IMDT stream that lives in pieces (ref aux_separator.h:627-656 is one sequential memmem
    loop): a rank keeps the signatures that start inside its piece (found in piece + halo), all ranks exchange (global offset,
    trailer) -- a few hundred entries of 180 bytes, one all_gather_object -- and each runs the host chain over the whole
    table, so every rank ends up with the same frame table, offsets relative to the whole stream.
    Returns (entries, stats[4])."""
    world = len(piece_bytes)
    payload = frames_local_hits(find_hits, piece_bytes, rank)
    if world > 1:
        import torch.distributed as dist
        allp = [None] * world
        dist.all_gather_object(allp, payload, group=group)
    else:
        allp = [payload]
    return frames_chain_all(allp, piece_bytes, tile_cols, tile_lines)
