"""Host-side Python mirror of the reference's per-pixel functions over the C ABI.

PyTorch is used for device memory and streams only; every operation below is one or more calls
into liboip_b200.so (hand-written sm_100a kernels).  Names follow the reference:

    inplace_rrc            IMO::InplaceRRC                      ref imageop.h:129-138
    prestitch_shift        Stitcher::PreStitch/SectionaryRemap  ref stitcher.h:83-139, imageop.h:230-275
    stitch_big_raw         IMO::StitchBigRaw                    ref imageop.h:277-363
    pan_pipeline           the three above fused, N CCDs
    band_align             PreProcessor::DoInterBandAlignment   ref preproc.h:351-468
    stitch_tiff_geometry   IMO::StitchTiff* geometry            ref imageop.h:416-421, :501-538
    aos_scan / imtr_deframe / image_frames_index / unpack_frames
                           AuxSeparator                         ref aux_separator.h:256-690
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import capi
from .capi import (FMT_BE16, FMT_BE16_TILES, FMT_LE16, FMT_PACK10, FMT_PACK12, CcdSrc, FrameEntry, FrameGeom,
                   MssDesc, PanDesc, RowSeg, check)

SECTION_ROWS = 30000  # REMAP_SECTION_ROWS, ref imageop.h:20
ROW_GUARD = 32767     # REMAP_ROW_GUARD,    ref imageop.h:19


class Context:
    """one per GPU; owns a CUDA stream unless bound to torch's current stream."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        self.lib = capi.load()
        self.device = device
        h = C.c_void_p()
        stream = None
        if use_torch_stream:
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        check(self.lib.oip_ctx_create(device, stream, 0 if use_torch_stream else 1, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.oip_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.oip_ctx_sync(self.h))

    def set_option(self, name: str, value: int):
        check(self.lib.oip_ctx_set_option(self.h, name.encode(), int(value)))

    @property
    def launches(self) -> int:
        return int(self.lib.oip_ctx_launch_count(self.h))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def u16(t: torch.Tensor) -> torch.Tensor:
    return t.view(torch.uint16) if t.dtype != torch.uint16 else t


# ------------------------------------------------------------------------------- stage 2 / 3 (PAN)
def make_pan_desc(ccds: Sequence[torch.Tensor], fmt, kbs, dX, dY, shifted, fold_half: int, out: torch.Tensor,
                  total_rows: Optional[int] = None, row0: int = 0, n_rows: Optional[int] = None,
                  section_rows: int = SECTION_ROWS, row_guard: int = ROW_GUARD, segs=None, pitch_bytes=None,
                  w: Optional[int] = None) -> PanDesc:
    """ccds[i]: device tensor holding rows of CCD i (2-D u16 for 16-bit formats, 2-D u8 for packed).
    segs[i] (optional): list of (tensor_or_ptr, row0, n_rows, pitch_bytes) overriding the single segment."""
    d = PanDesc()
    n = len(ccds)
    d.n_ccd = n
    if w is None:
        w = ccds[0].shape[1]
    d.w = w
    h = ccds[0].shape[0]
    d.total_rows = h if total_rows is None else total_rows
    d.row0 = row0
    d.n_rows = (d.total_rows - row0) if n_rows is None else n_rows
    d.fold_half = fold_half
    d.section_rows = section_rows
    d.row_guard = row_guard
    for i in range(n):
        c = d.ccd[i]
        c.fmt = fmt if isinstance(fmt, int) else fmt[i]
        if segs is not None and segs[i] is not None:
            c.n_seg = len(segs[i])
            for s, (base, r0, nr, pb) in enumerate(segs[i]):
                c.seg[s] = RowSeg(base if isinstance(base, int) else base.data_ptr(), r0, nr, pb)
        else:
            c.n_seg = 1
            pb = ccds[i].stride(0) * ccds[i].element_size() if pitch_bytes is None else pitch_bytes
            c.seg[0] = RowSeg(ccds[i].data_ptr(), 0, ccds[i].shape[0], pb)
        c.d_kb = _ptr(kbs[i]) if kbs is not None and kbs[i] is not None else None
        c.shifted = int(bool(shifted[i]))
        c.dX = float(dX[i])
        c.dY = float(dY[i])
    d.d_out = out.data_ptr()
    d.out_pitch_px = out.stride(0)
    return d


def pan_out_width(n_ccd: int, w: int, fold_half: int) -> int:
    return capi.load().oip_pan_out_width(n_ccd, w, fold_half)


def pan_pipeline(ctx: Context, ccds, kbs, dX, dY, fold_half: int, fmt=FMT_LE16, shifted=None,
                 section_rows: int = SECTION_ROWS, row_guard: int = ROW_GUARD, out: Optional[torch.Tensor] = None,
                 check_error: bool = True, w: Optional[int] = None) -> torch.Tensor:
    """fused unpack -> RRC -> shift -> concat.  CCD 0 is copied, CCD i>=1 shifted by (dX[i], dY[i])."""
    n = len(ccds)
    if w is None:
        w = ccds[0].shape[1]
    h = ccds[0].shape[0]
    if shifted is None:
        shifted = [i > 0 for i in range(n)]
    if out is None:
        out = torch.empty((h, pan_out_width(n, w, fold_half)), dtype=torch.uint16, device=ccds[0].device)
    d = make_pan_desc(ccds, fmt, kbs, dX, dY, shifted, fold_half, out, section_rows=section_rows,
                      row_guard=row_guard, w=w)
    check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
    if check_error:
        check(ctx.lib.oip_pan_check_error(ctx.h))
    return out


def pan_pipeline_host(ctx: Context, ccds_host, kbs_host, dX, dY, fold_half: int, out_host: torch.Tensor,
                      fmt=FMT_LE16, shifted=None, section_rows: int = SECTION_ROWS, row_guard: int = ROW_GUARD,
                      w: Optional[int] = None) -> torch.Tensor:
    """same operation for HOST buffers (ideally pinned): host->device, kernels and device->host are
    overlapped in row blocks inside the library (oip_pan_pipeline_host).  Returns out_host."""
    n = len(ccds_host)
    if shifted is None:
        shifted = [i > 0 for i in range(n)]
    d = make_pan_desc(ccds_host, fmt, kbs_host, dX, dY, shifted, fold_half, out_host, section_rows=section_rows,
                      row_guard=row_guard, w=w)
    check(ctx.lib.oip_pan_pipeline_host(ctx.h, C.byref(d)))
    return out_host


def frame_tile_table(ents, n_frames: int) -> np.ndarray:
    """(n_frames, 40) int64 sub-image byte offsets from oip_image_frames_index entries (-1 rows = zero-filled gap frames)"""
    return np.array([[ents[k].tile_off[j] for j in range(40)] for k in range(n_frames)], np.int64).reshape(n_frames, 40)


def pan_pipeline_from_frames(ctx: Context, imdts, tables, tile_cols: int, tile_lines: int, kbs, dX, dY, fold_half: int,
                             shifted=None, section_rows: int = SECTION_ROWS, row_guard: int = ROW_GUARD,
                             out: Optional[torch.Tensor] = None, check_error: bool = True, keep=None):
    """K9: the fused PAN path straight from the image frames of the IMDT streams -- no PAN raster is
    written in between (the reference writes .PAN.RAW, ref aux_separator.h:341-372, then .RRC.RAW and .PRESTT.RAW).
    imdts[i]: device uint8 IMDT stream of CCD i; tables[i]: frame_tile_table() of its frames (host int64 [n_frames, 40])."""
    n = len(imdts)
    n_frames = min(t.shape[0] for t in tables)
    rows = n_frames * 4 * tile_lines
    w = 8 * tile_cols
    if shifted is None:
        shifted = [i > 0 for i in range(n)]
    if out is None:
        out = torch.empty((rows, pan_out_width(n, w, fold_half)), dtype=torch.uint16, device=imdts[0].device)
    d = PanDesc()
    d.n_ccd, d.w, d.total_rows, d.row0, d.n_rows = n, w, rows, 0, rows
    d.fold_half, d.section_rows, d.row_guard = fold_half, section_rows, row_guard
    hold = [] if keep is None else keep
    for i in range(n):
        c = d.ccd[i]
        c.fmt, c.n_seg = FMT_BE16_TILES, 1
        c.seg[0] = RowSeg(imdts[i].data_ptr(), 0, rows, 0)
        c.d_kb = _ptr(kbs[i]) if kbs is not None and kbs[i] is not None else None
        c.shifted, c.dX, c.dY = int(bool(shifted[i])), float(dX[i]), float(dY[i])
        h_tab = np.ascontiguousarray(tables[i][:n_frames], np.int64)
        d_tab = torch.from_numpy(h_tab).to(imdts[i].device)
        hold += [h_tab, d_tab]
        c.d_tile_off, c.h_tile_off, c.tile_cols, c.tile_lines = d_tab.data_ptr(), h_tab.ctypes.data, tile_cols, tile_lines
    d.d_out, d.out_pitch_px = out.data_ptr(), out.stride(0)
    check(ctx.lib.oip_pan_pipeline(ctx.h, C.byref(d)))
    if check_error or keep is None:
        check(ctx.lib.oip_pan_check_error(ctx.h))   # synchronises: the tables may be released afterwards
    return out, d


def downlink_to_stitched(ctx: Context, files, tile_cols: int, tile_lines: int, kbs, dX, dY, fold_half: int, shifted=None,
                         section_rows: int = SECTION_ROWS, row_guard: int = ROW_GUARD, out: Optional[torch.Tensor] = None,
                         want_aux: bool = False, want_mss: bool = False):
    """raw downlink files (device uint8 tensors, one per CCD) -> stitched PAN raster in ONE call: AOS scan + CRC, IMTR
    re-framing, image-frame index, then the fused RRC + shift + stitch reading the sub-images where they lie
    (oip_downlink_to_stitched).  Returns (out[:rows], stats list, aux list, mss list)."""
    from .capi import DownlinkDesc, DownlinkStats
    n = len(files)
    w = 8 * tile_cols
    if shifted is None:
        shifted = [i > 0 for i in range(n)]
    # lines: bounded by the payload the smallest file can carry
    frame_bytes = 192 * tile_lines + 40 * tile_lines * tile_cols * 2 + 172
    cap_rows = (min(f.numel() for f in files) // frame_bytes + 1) * 4 * tile_lines
    if out is None:
        out = torch.empty((cap_rows, pan_out_width(n, w, fold_half)), dtype=torch.uint16, device=files[0].device)
    d = DownlinkDesc()
    d.n_ccd, d.geom = n, FrameGeom(tile_cols, tile_lines)
    d.fold_half, d.section_rows, d.row_guard = fold_half, section_rows, row_guard
    aux, mss = [None] * n, [None] * n
    for i in range(n):
        c = d.ccd[i]
        c.d_file, c.n_bytes = files[i].data_ptr(), files[i].numel()
        c.d_kb = _ptr(kbs[i]) if kbs is not None and kbs[i] is not None else None
        c.shifted, c.dX, c.dY = int(bool(shifted[i])), float(dX[i]), float(dY[i])
        if want_aux:
            aux[i] = torch.zeros((cap_rows // (4 * tile_lines), 192 * tile_lines), dtype=torch.uint8, device=files[i].device)
            d.d_aux[i] = aux[i].data_ptr()
        if want_mss:
            mss[i] = torch.zeros((cap_rows // 4, w), dtype=torch.uint16, device=files[i].device)
            d.d_mss[i] = mss[i].data_ptr()
    d.d_out, d.out_pitch_px, d.out_rows_cap = out.data_ptr(), out.stride(0), out.shape[0]
    rows = C.c_int64(0)
    st = (DownlinkStats * n)()
    check(ctx.lib.oip_downlink_to_stitched(ctx.h, C.byref(d), C.byref(rows), st))
    r = int(rows.value)
    nf = r // (4 * tile_lines)
    stats = [dict(aos=list(s.aos), imtr=list(s.imtr), frames=list(s.frames), imdt_bytes=int(s.imdt_bytes)) for s in st]
    return out[:r], stats, [a[:nf] if a is not None else None for a in aux], [m[:nf * tile_lines] if m is not None else None for m in mss]


def inplace_rrc(ctx: Context, img: torch.Tensor, kb: torch.Tensor) -> torch.Tensor:
    h, w = img.shape
    check(ctx.lib.oip_rrc_u16(ctx.h, img.data_ptr(), w, h, img.stride(0), kb.data_ptr()))
    return img


def prestitch_shift(ctx: Context, src: torch.Tensor, dX: float, dY: float, section_rows: int = SECTION_ROWS,
                    row_guard: int = ROW_GUARD) -> torch.Tensor:
    h, w = src.shape
    dst = torch.empty_like(src)
    check(ctx.lib.oip_shift_cubic_u16(ctx.h, src.data_ptr(), dst.data_ptr(), w, h, dX, dY, section_rows, row_guard))
    check(ctx.lib.oip_pan_check_error(ctx.h))
    return dst


def stitch_big_raw(ctx: Context, ccds: Sequence[torch.Tensor], fold_half: int) -> torch.Tensor:
    n = len(ccds)
    h, w = ccds[0].shape
    out = torch.empty((h, pan_out_width(n, w, fold_half)), dtype=torch.uint16, device=ccds[0].device)
    arr = (C.c_void_p * n)(*[c.data_ptr() for c in ccds])
    check(ctx.lib.oip_stitch_concat_u16(ctx.h, arr, n, w, h, fold_half, out.data_ptr()))
    check(ctx.lib.oip_pan_check_error(ctx.h))
    return out


def cubic_tab() -> np.ndarray:
    t = np.zeros(128, np.float32)
    capi.load().oip_cubic_tab(t.ctypes.data)
    return t.reshape(32, 4)


# ------------------------------------------------------------------------------- stage 3 (MSS)
def band_align(ctx: Context, mss: torch.Tensor, wb: int, kbs, cX, cY, lines_per_section: int = 20000,
               line_offset: int = 0, overlap: int = 520, keep_leading: bool = False, min_process_lines: int = 1500,
               fmt: int = FMT_LE16, out: Optional[torch.Tensor] = None, total_lines: Optional[int] = None, src_row0: int = 0,
               sec_first: int = 0, sec_count: int = 0):
    """sec_count > 0: section shard -- mss holds strip lines [src_row0, src_row0 + len(mss)) of a total_lines strip and
    `out` row 0 is the first output row of section sec_first (sharding.mss_sections / mss_rank_sections)"""
    lines = mss.shape[0] if total_lines is None else int(total_lines)
    d = MssDesc()
    d.fmt = fmt
    d.wb = wb
    d.lines = lines
    d.pitch_px = mss.stride(0)
    for b in range(4):
        d.d_kb[b] = _ptr(kbs[b]) if kbs is not None and kbs[b] is not None else None
    cX = np.asarray(cX, np.float64).reshape(-1)
    cY = np.asarray(cY, np.float64).reshape(-1)
    for i in range(8):
        d.cX[i] = cX[i]
    for i in range(12):
        d.cY[i] = cY[i]
    d.lines_per_section = lines_per_section
    d.line_offset = line_offset
    d.overlap = overlap
    d.keep_leading = int(keep_leading)
    d.min_process_lines = min_process_lines
    d.sec_first, d.sec_count, d.src_row0 = sec_first, sec_count, src_row0
    rows = lines - line_offset - (0 if keep_leading else overlap)
    if out is None:
        out = torch.zeros((max(rows, 0), wb, 4), dtype=torch.uint16, device=mss.device)
    n = C.c_int64(0)
    check(ctx.lib.oip_band_align_merge(ctx.h, mss.data_ptr(), C.byref(d), out.data_ptr(), C.byref(n)))
    return int(n.value), out


def band_align_sections(ctx: Context, mss: torch.Tensor, wb: int, kbs, cX, cY, sections, all_sections, out: torch.Tensor,
                        total_lines: int, lines_per_section: int = 20000, overlap: int = 520, line_offset: int = 0,
                        keep_leading: bool = False, min_process_lines: int = 1500, fmt: int = FMT_LE16, src_row0: int = 0) -> int:
    """the band alignment of a contiguous run of sections (sharding.mss_rank_sections) in ONE call: a section depends only on
    its own source lines, so a rank that holds lines [src_row0, src_row0 + len(mss)) produces the output rows of its
    sections bit-identical to the whole-strip call.  `out` row 0 = first output row of the first section given."""
    if not sections:
        return 0
    first = all_sections.index(sections[0])
    assert list(all_sections[first:first + len(sections)]) == list(sections), "sections must be a contiguous run"
    assert sections[0][0] >= src_row0 and sections[-1][0] + sections[-1][1] <= src_row0 + mss.shape[0], \
        "the shard does not hold its sections' source lines"
    n, _ = band_align(ctx, mss, wb, kbs, cX, cY, lines_per_section=lines_per_section, line_offset=line_offset, overlap=overlap,
                      keep_leading=keep_leading, min_process_lines=min_process_lines, fmt=fmt, out=out, total_lines=total_lines,
                      src_row0=src_row0, sec_first=first, sec_count=len(sections))
    return n


def stitch_tiff_geometry(ctx: Context, imgs: Sequence[torch.Tensor], fold_half: int, band_map=None) -> torch.Tensor:
    n = len(imgs)
    h, w, _ = imgs[0].shape
    out = torch.empty((h, pan_out_width(n, w, fold_half), 4), dtype=torch.uint16, device=imgs[0].device)
    arr = (C.c_void_p * n)(*[c.data_ptr() for c in imgs])
    bm = (C.c_int * 4)(*band_map) if band_map is not None else None
    check(ctx.lib.oip_stitch_concat_c4(ctx.h, arr, n, w, h, fold_half, bm, out.data_ptr()))
    return out


def unpack_lines(ctx: Context, raw: torch.Tensor, fmt: int, w: int) -> torch.Tensor:
    rows = raw.shape[0]
    out = torch.empty((rows, w), dtype=torch.uint16, device=raw.device)
    check(ctx.lib.oip_unpack_lines(ctx.h, raw.data_ptr(), fmt, w, rows, raw.stride(0) * raw.element_size(),
                                   out.data_ptr()))
    return out


# ------------------------------------------------------------------------------- stage 1
def crc16_batch(ctx: Context, buf: torch.Tensor, off: torch.Tensor, length: int) -> torch.Tensor:
    out = torch.empty(off.numel(), dtype=torch.uint16, device=buf.device)
    check(ctx.lib.oip_crc16_batch(ctx.h, buf.data_ptr(), off.data_ptr(), off.numel(), length, out.data_ptr()))
    return out


def aos_scan(ctx: Context, buf: torch.Tensor):
    """returns (payload_off device int64 tensor [n_valid], counters np.int64[3] = valid, invalid, empty)"""
    n = buf.numel()
    cap = n // 1024 + 1
    off = torch.empty(cap, dtype=torch.int64, device=buf.device)
    cnt = (C.c_int64 * 3)()
    check(ctx.lib.oip_aos_scan(ctx.h, buf.data_ptr(), n, off.data_ptr(), cap, cnt))
    return off[: cnt[0]], np.array(list(cnt), np.int64)


def aos_scan_shard(ctx: Context, buf: torch.Tensor, own_bytes: int, carry_in: int = 0):
    """byte-range shard of the AOS scan (oip_aos_scan_shard): buf = the shard's own bytes + up to 1023 halo bytes.
    Returns (payload_off relative to buf [n_valid], counters np.int64[3], carry_out)."""
    n = buf.numel()
    cap = n // 1024 + 1
    off = torch.empty(cap, dtype=torch.int64, device=buf.device)
    cnt = (C.c_int64 * 3)()
    co = C.c_int64(0)
    check(ctx.lib.oip_aos_scan_shard(ctx.h, buf.data_ptr(), n, own_bytes, carry_in, off.data_ptr(), cap, cnt, C.byref(co)))
    return off[: cnt[0]], np.array(list(cnt), np.int64), int(co.value)


def imtr_deframe_shard(ctx: Context, buf: torch.Tensor, payload_off: torch.Tensor, skip: int, n_frames: int, prev_seq: int = -1):
    """shard of the IMTR re-framing (oip_imtr_deframe_shard).  Returns (imdt piece, info dict for sharding.imtr_combine)."""
    n = payload_off.numel()
    cap = (n_frames + 1) * 866
    imdt = torch.empty(cap, dtype=torch.uint8, device=buf.device)
    st = (C.c_int64 * 9)()
    nb = C.c_int64(0)
    si = (C.c_int64 * 3)()
    check(ctx.lib.oip_imtr_deframe_shard(ctx.h, buf.data_ptr(), payload_off.data_ptr(), n, skip, n_frames, prev_seq, imdt.data_ptr(), cap,
                                         st, C.byref(nb), si))
    info = dict(n_frames=int(st[0]), n_valid=int(st[1]), bad=[int(st[2]), int(st[3]), int(st[4]), int(st[5])], first_seq=int(si[0]),
                last_seq=int(si[1]), gaps=int(st[6]), restarts=int(st[8]), local_restart=int(si[2]), first_chid=int(st[7]),
                imdt_bytes=int(nb.value))
    return imdt[: nb.value], info


def imtr_deframe(ctx: Context, buf: torch.Tensor, payload_off: torch.Tensor):
    n = payload_off.numel()
    cap = (n * 880 // 882 + 1) * 866
    imdt = torch.empty(cap, dtype=torch.uint8, device=buf.device)
    st = (C.c_int64 * 9)()
    nb = C.c_int64(0)
    check(ctx.lib.oip_imtr_deframe(ctx.h, buf.data_ptr(), payload_off.data_ptr(), n, imdt.data_ptr(), cap, st,
                                   C.byref(nb)))
    return imdt[: nb.value], np.array(list(st), np.int64)


def image_frames_index(ctx: Context, imdt: torch.Tensor, tile_cols: int, tile_lines: int):
    g = FrameGeom(tile_cols, tile_lines)
    st = (C.c_int64 * 4)()
    # ONE pass: the table is sized for every complete frame the stream can hold (+ slack for zero-filled gap frames);
    # only a stream with long sequence gaps needs the second call
    frame_bytes = 192 * tile_lines + 40 * tile_lines * tile_cols * 2 + 172
    cap = imdt.numel() // frame_bytes + 64
    ents = (FrameEntry * cap)()
    rc = ctx.lib.oip_image_frames_index(ctx.h, imdt.data_ptr(), imdt.numel(), C.byref(g), ents, cap, st)
    if rc == capi.OIP_E_INVALID and st[1] > cap:
        cap = int(st[1])
        ents = (FrameEntry * cap)()
        rc = ctx.lib.oip_image_frames_index(ctx.h, imdt.data_ptr(), imdt.numel(), C.byref(g), ents, cap, st)
    check(rc)
    return ents, np.array(list(st), np.int64)


def image_frames_hits(ctx: Context, imdt: torch.Tensor):
    """device half of the frame index (oip_image_frames_hits): ascending offsets of the trailer signature in `imdt` and the
    172 bytes that start at each (zero padded past the end) -> (uint64[n], uint8[n, 172]) on the host"""
    cap = max(64, imdt.numel() // 4096 + 64)
    for _ in range(2):
        hits = np.zeros(cap, np.uint64)
        tr = np.zeros((cap, 172), np.uint8)
        n = C.c_int64(0)
        rc = ctx.lib.oip_image_frames_hits(ctx.h, imdt.data_ptr(), imdt.numel(), hits.ctypes.data, tr.ctypes.data, cap, C.byref(n))
        if rc == capi.OIP_E_INVALID and n.value > cap:
            cap = int(n.value)
            continue
        check(rc)
        return hits[: n.value].copy(), tr[: n.value].copy()
    raise RuntimeError("oip_image_frames_hits: capacity did not settle")


def image_frames_chain(hits: np.ndarray, trailers: np.ndarray, n_bytes: int, tile_cols: int, tile_lines: int):
    """host half of the frame index (oip_image_frames_chain; no GPU, no context): the reference's frame chain and gap rules
    over a signature table with offsets relative to the whole IMDT stream of n_bytes -> (entries, stats[4])"""
    lib = capi.load()
    hits = np.ascontiguousarray(hits, np.uint64)
    trailers = np.ascontiguousarray(trailers, np.uint8).reshape(-1, 172) if len(hits) else np.zeros((0, 172), np.uint8)
    assert trailers.shape[0] == hits.shape[0]
    g = FrameGeom(tile_cols, tile_lines)
    st = (C.c_int64 * 4)()
    frame_bytes = 192 * tile_lines + 40 * tile_lines * tile_cols * 2 + 172
    cap = int(n_bytes) // frame_bytes + 64
    for _ in range(2):
        ents = (FrameEntry * cap)()
        rc = lib.oip_image_frames_chain(hits.ctypes.data if len(hits) else None, trailers.ctypes.data if len(hits) else None, len(hits),
                                        int(n_bytes), C.byref(g), ents, cap, st)
        if rc == capi.OIP_E_INVALID and st[1] > cap:
            cap = int(st[1])
            continue
        check(rc)
        return ents, np.array(list(st), np.int64)
    raise RuntimeError("oip_image_frames_chain: capacity did not settle")


def unpack_frames(ctx: Context, imdt: torch.Tensor, tile_cols: int, tile_lines: int, ents, n_frames: int,
                  want_aux=True, want_pan=True, want_mss=True):
    g = FrameGeom(tile_cols, tile_lines)
    W = 8 * tile_cols
    dev = imdt.device
    aux = torch.empty((n_frames, 192 * tile_lines), dtype=torch.uint8, device=dev) if want_aux else None
    pan = torch.empty((n_frames * 4 * tile_lines, W), dtype=torch.uint16, device=dev) if want_pan else None
    mss = torch.empty((n_frames * tile_lines, W), dtype=torch.uint16, device=dev) if want_mss else None
    check(ctx.lib.oip_unpack_frames(ctx.h, imdt.data_ptr(), imdt.numel(), C.byref(g), ents, n_frames, _ptr(aux),
                                    _ptr(pan), _ptr(mss)))
    return aux, pan, mss


# ------------------------------------------------------------------------------------------------
# inter-CMOS offset estimation (ref stitcher.h:148-201)
# ------------------------------------------------------------------------------------------------
def phase_correlate(ctx: Context, a: torch.Tensor, b: torch.Tensor):
    """cv::phaseCorrelate(a, b) on two u16 device images (row views allowed) -> (dx, dy, response)"""
    assert a.shape == b.shape and a.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    res = (C.c_double * 3)()
    check(ctx.lib.oip_phase_correlate_u16(ctx.h, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), a.shape[0], a.shape[1], res))
    return res[0], res[1], res[2]


def calc_stt_parameters(ctx: Context, pan1: torch.Tensor, pan2: torch.Tensor, overlap_cols: int = 200, edge_cols: int = 0,
                        sections: int = 10, lines_per_section: int = 16000, threshold: float = 0.4, max_delta_y: float = 0.0,
                        total_lines: Optional[int] = None, row0: int = 0, group=None):
    """Stitcher::CalcSttParameters: per-section rows (line_offset, dx, dy, response, valid) and the mean
    (dx, dy, response) over the valid sections, or None when there is none (the reference throws there).
    pan1 / pan2 may be one scanline-block shard (rows [row0, row0+len) of a total_lines strip): sections the shard
    holds entirely are correlated here and the four sums are added over `group` with ONE small all-reduce."""
    from .capi import SttConfig, SttSection
    assert pan1.shape == pan2.shape and pan1.stride(1) == 1 and pan1.stride(0) == pan2.stride(0)
    rows_here, w = pan1.shape
    total = rows_here if total_lines is None else int(total_lines)
    cfg = SttConfig(sections, lines_per_section, overlap_cols, edge_cols, threshold, max_delta_y)
    secs = (SttSection * sections)()
    sums = (C.c_double * 4)()
    check(ctx.lib.oip_stt_parameters(ctx.h, pan1.data_ptr(), pan2.data_ptr(), w, total, row0, rows_here, pan1.stride(0),
                                     C.byref(cfg), secs, sums))
    rows = [(s.line_offset, s.dx, s.dy, s.response, s.valid) for s in secs]
    tot = [sums[0], sums[1], sums[2], sums[3]]
    if group is not None or (total_lines is not None and torch.distributed.is_available() and torch.distributed.is_initialized()):
        from . import sharding
        dist = torch.distributed
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        # sections that straddle two scanline blocks: their rows of the two overlap slices go to ONE rank (a few MB,
        # point to point), which correlates them like any other section (ref stitcher.h:166-199 skips none)
        mine = torch.tensor([row0, row0 + rows_here], dtype=torch.int64, device=pan1.device)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine, group=group)
        plan = sharding.stt_section_plan(total, sections, lines_per_section, [tuple(t.tolist()) for t in allr])
        for idx, a, b in sharding.stt_gather_straddling(pan1, pan2, row0, rank, plan, (w - overlap_cols, w - edge_cols),
                                                        (edge_cols, overlap_cols), group):
            dx, dy, rs = phase_correlate(ctx, a, b)
            ok = rs >= threshold and (max_delta_y <= 0.0 or abs(dy) <= max_delta_y)          # ref stitcher.h:181
            rows[idx] = (rows[idx][0], dx, dy, rs, int(ok))
            if ok:
                tot = [tot[0] + dx, tot[1] + dy, tot[2] + rs, tot[3] + 1.0]
        t = torch.tensor(tot, dtype=torch.float64, device=pan1.device)
        dist.all_reduce(t, group=group)
        tot = t.tolist()
    mean = None if tot[3] == 0 else (tot[0] / tot[3], tot[1] / tot[3], tot[2] / tot[3])
    return rows, mean


# ------------------------------------------------------------------------------------------------
# inter-band shift estimation + polynomial fit (ref preproc.h:224-347, :492-550)
# ------------------------------------------------------------------------------------------------
def calc_inter_band_correlation(ctx: Context, pan: torch.Tensor, mss: torch.Tensor, slices: int = 10, sections: int = 5,
                                threshold: float = 0.4, correlation_lines: int = 16000, min_slices: int = 8, min_count: int = 5):
    """PreProcessor::CalcInterBandCorrelation: pan (lines x W) and mss (lines/4 x W, 4 bands side by side) on the device
    -> (shifts[band][sec*slices+i] = (dx, dy, rs, cx), cX[4][2], cY[4][3])"""
    from .capi import IbcConfig, IbcShift
    assert pan.stride(1) == 1 and mss.stride(1) == 1 and pan.shape[1] == mss.shape[1]
    cfg = IbcConfig(slices, sections, threshold, correlation_lines, min_slices, min_count, 0)
    n = slices * sections
    sh = (IbcShift * (4 * n))()
    cX, cY = (C.c_double * 8)(), (C.c_double * 12)()
    check(ctx.lib.oip_inter_band_correlation(ctx.h, pan.data_ptr(), pan.shape[1], pan.shape[0], pan.stride(0), mss.data_ptr(),
                                             mss.shape[0], mss.stride(0), C.byref(cfg), sh, cX, cY))
    shifts = [[(sh[b * n + k].dx, sh[b * n + k].dy, sh[b * n + k].rs, sh[b * n + k].cx) for k in range(n)] for b in range(4)]
    return shifts, [[cX[2 * b], cX[2 * b + 1]] for b in range(4)], [[cY[3 * b], cY[3 * b + 1], cY[3 * b + 2]] for b in range(4)]
