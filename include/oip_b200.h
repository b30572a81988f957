/*
 * oip_b200.h -- C ABI of the B200-native OpticalImageProcessor hot path (liboip_b200.so).
 *
 * The reference (arloan/OpticalImageProcessor) has no plugin/FFI layer: its per-pixel path is a
 * set of C++ static/member functions called from main.cpp.  Each entry point below replaces one
 * of those functions and cites it as  ref <file>:<line>  (= /root/reference/OpticalImageProcessor/).
 * INTEGRATION.md shows the call-site patch a reference maintainer would apply.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No C++/torch types.  No exception crosses this boundary.
 *   - every function returns OIP_OK (0) or a negative oip_status; oip_last_error() gives the text
 *     (thread-local).  The reference throws std::invalid_argument / std::runtime_error at the same
 *     conditions; the CLI maps a negative status to the reference's exit code 2 (ref main.cpp:336-338).
 *   - "d_" pointers are device pointers on the context's device, caller-owned; the library never
 *     frees caller memory.  Work is enqueued on the context's stream; functions that return host
 *     scalars (counters, sizes) synchronise that stream before returning, the others do not.
 *   - pixels are uint16 little-endian, rows are `pitch_px` pixels apart (ref oipshared.h:27-29).
 *   - there is NO CPU fallback: without a usable sm_100 device every call fails with OIP_E_CUDA.
 */
#ifndef OIP_B200_H
#define OIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIP_ABI_VERSION 2

typedef enum {
    OIP_OK = 0,
    OIP_E_INVALID = -1,  /* bad argument (the reference throws std::invalid_argument) */
    OIP_E_CUDA = -2,     /* CUDA runtime error / no device */
    OIP_E_NOMEM = -3,
    OIP_E_RANGE = -4,    /* a frame/tile read would leave the input buffer */
    OIP_E_UNSUPPORTED = -5, /* e.g. JPEG-2000 compressed sub-images (ref aux_separator.h:378-383) */
    OIP_E_IO = -6
} oip_status;

typedef struct oip_ctx oip_ctx;

/* ---- context ----------------------------------------------------------------------------- */
/* own_stream != 0: the context creates its own non-blocking stream (`stream` ignored);
 * own_stream == 0: enqueue on `stream`, a cudaStream_t (0 = the legacy default stream, which is
 * what torch's current stream is unless the caller changed it) */
int oip_ctx_create(int device, void *stream, int own_stream, oip_ctx **out);
void oip_ctx_destroy(oip_ctx *ctx);
int oip_ctx_sync(oip_ctx *ctx);
void *oip_ctx_stream(oip_ctx *ctx);
const char *oip_last_error(void);
int oip_abi_version(void);
/* number of kernels this library launched on ctx since creation (bench.py "gpu_launches") */
int64_t oip_ctx_launch_count(oip_ctx *ctx);

/* tunables / test switches.  "pan_fast" (0|1, default 1): 0 sends every PAN tile through the generic kernel;
 * "pan_fast_stages" (2..8): TMA stages per warp; "pan_fast_rows": output rows per warp-tile;
 * "mss_fast" (0|1), "mss_fast_rows": the same switches for oip_band_align_merge;
 * "aos_fused" (0|1, default 1): 0 = oip_aos_scan searches every byte (aos_scan_kernel + aos_crc_kernel) instead of
 * the single cadence pass; "imtr_runs" (0|1, default 1): 0 = oip_imtr_deframe gathers frame by frame instead of run by
 * run; "downlink_threads" (0|1, default 1): 0 = oip_downlink_to_stitched runs stage 1 of its CCDs one after the other
 * on the caller's stream instead of side by side; "host_block_rows": rows per block of oip_pan_pipeline_host.
 * Every switch selects between two implementations with identical results (the tests run both). */
int oip_ctx_set_option(oip_ctx *ctx, const char *name, int64_t value);

/* raw memory helpers for hosts without their own allocator (the CLI); torch callers pass data_ptr() */
int oip_dev_alloc(oip_ctx *ctx, size_t bytes, void **d_ptr);
int oip_dev_free(oip_ctx *ctx, void *d_ptr);
int oip_host_alloc_pinned(size_t bytes, void **h_ptr);
int oip_host_free_pinned(void *h_ptr);
int oip_copy_h2d(oip_ctx *ctx, void *d_dst, const void *h_src, size_t bytes); /* async on ctx stream */
int oip_copy_d2h(oip_ctx *ctx, void *h_dst, const void *d_src, size_t bytes); /* async on ctx stream */
int oip_memset_d(oip_ctx *ctx, void *d_dst, int value, size_t bytes);

/* multi-GPU halo rows over NVLink P2P: export a device allocation to the neighbour ranks
 * (cudaIpcMemHandle_t is 64 bytes) and map a neighbour's.  Used by bench.py / tests at N>1. */
int oip_ipc_export(oip_ctx *ctx, void *d_ptr, uint8_t handle[64]);
int oip_ipc_open(oip_ctx *ctx, const uint8_t handle[64], void **d_peer_ptr);
int oip_ipc_close(oip_ctx *ctx, void *d_peer_ptr);

/* ---- stage 1: frame handling -------------------------------------------------------------- */

/* CRC-16/CCITT-FALSE of n_items byte ranges [d_off[i], d_off[i]+len) of d_buf.
 * replaces CRC::Calculate(data, size, CRC::CRC_16_CCITTFALSE()) -- ref CRC.h:456-464 as called at
 * aux_separator.h:579 and :679-681. */
int oip_crc16_batch(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_off, int64_t n_items,
                    int len, uint16_t *d_crc);

/* The chained AOS scan: sync search + ValidateAosFrame + skip rules.
 * replaces AuxSeparator::SeparateAosFile / NextAosFrame / ValidateAosFrame --
 * ref aux_separator.h:395-467, :622-625, :658-690.
 * d_payload_off[i] (capacity cap, >= n_bytes/1024+1 is always enough) receives the byte offset of
 * the 880-byte payload of the i-th valid frame, in file order.  counters = {valid, invalid, empty}
 * (ref :411-413).  A sync hit counts only if off+1024 <= n_bytes. */
int oip_aos_scan(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, uint64_t *d_payload_off,
                 size_t cap, int64_t counters[3]);

/* Byte-range shard of the same scan (the downlink of one strip spread over the GPUs of a box).  d_buf holds
 * file bytes [B, B + n_bytes): the shard owns the candidates that start in its first own_bytes, the rest (up to 1023
 * bytes, read from the file by the same rank) is the halo that lets the frames starting near the end be validated.
 * carry_in = buffer offset where this shard's scan starts (0, or the bytes of the previous shard's last accepted frame that
 * reach in here); *carry_out = the same quantity for the next shard.  Shards are independent given carry_in; ranks
 * exchange (carry_out, n_valid) with ONE all-gather, re-run the rare shard whose assumed carry_in was wrong, and add the
 * three counters with one all-reduce (opticalimageprocessor_b200/sharding.py: aos_shard_ranges / aos_resolve_carries).
 * Payload offsets are relative to d_buf.  Concatenating the shards' results gives oip_aos_scan of the whole file. */
int oip_aos_scan_shard(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, size_t own_bytes, size_t carry_in,
                       uint64_t *d_payload_off, size_t cap, int64_t counters[3], int64_t *carry_out);

/* IMTR re-framing at the fixed 882-byte cadence over the concatenated payloads, validation and
 * extraction of the 866-byte bodies.
 * replaces AuxSeparator::DataTransFrameParser + ValidateImtrFrame -- ref aux_separator.h:469-590.
 * stats = {frames_cut, frames_valid, bad_sig, bad_endsig, bad_type, bad_crc, seq_gaps,
 *          first_chid, restarts}; *imdt_bytes = bytes written to d_imdt (capacity cap). */
int oip_imtr_deframe(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off,
                     int64_t n_payload, uint8_t *d_imdt, size_t cap, int64_t stats[9],
                     int64_t *imdt_bytes);

/* The same on a shard of the payload stream.  The cadence is cut from GLOBAL stream byte 0, so a rank whose
 * payloads are stream bytes [880 P, 880 (P + n_own)) owns the frames that START in that range: the first one begins
 * skip_bytes into its first payload, n_frames of them follow (sharding.imtr_shard_frames computes both from the all-gathered
 * payload counts); d_payload_off lists the rank's own payloads plus the 1-2 payloads of the next rank its last frame
 * reaches into.  prev_seq = sequence number accepted before this shard (0 at the start of the stream, < 0: not known yet
 * -- the restart / gap rule of the first valid frame is then left to the caller, who applies it after the all-gather of
 * seq_info).  seq_info = {first valid seq, last valid seq, index of the last restart among the valid frames or -1}.
 * stats and d_imdt are local to the shard; the IMDT stream is the concatenation of the ranks' pieces from the last
 * restart on (sharding.imtr_combine). */
int oip_imtr_deframe_shard(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off, int64_t n_payload,
                           int skip_bytes, int64_t n_frames, int64_t prev_seq, uint8_t *d_imdt, size_t cap, int64_t stats[9],
                           int64_t *imdt_bytes, int64_t seq_info[3]);

/* image-frame geometry, reference values in brackets (ref aux_separator.h:84-93) */
typedef struct {
    int tile_cols;  /* IMGSIG_IMBASE_COLS  [1536] : line = 8 * tile_cols px */
    int tile_lines; /* IMGSIG_IMBASE_LINES [256]  : frame = 4*tile_lines PAN + tile_lines MSS lines */
} oip_frame_geom;

/* one emitted image frame (real or zero-filled gap frame) */
typedef struct {
    int64_t frame_off;     /* byte offset of the aux block in the IMDT stream, -1 = zero-filled gap */
    int64_t tile_off[40];  /* byte offset of each sub-image r*8+c (ref aux_separator.h:347-356) */
    int32_t seq;
    int32_t z_ratio;
} oip_frame_entry;

/* Locate image frames: trailer-signature search on the device, chain + gap rules on the host.
 * replaces AuxSeparator::NextImageDataFrame and the loop of SeparateImageData --
 * ref aux_separator.h:627-656, :287-320.
 * entries (host array, capacity cap) receives the emitted frames in output order.
 * stats = {frames_found, frames_emitted, frames_incomplete, last_seq}. */
int oip_image_frames_index(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes,
                           const oip_frame_geom *geom, oip_frame_entry *entries, int64_t cap,
                           int64_t stats[4]);

/* The two halves of oip_image_frames_index, for an IMDT stream that lives in pieces on several GPUs: every
 * rank searches its own piece (plus the first 174 bytes of what follows, so that a signature or trailer cut by the piece
 * boundary is seen whole), the ranks all-gather the few hundred (offset, trailer) pairs, and each runs the chain over the
 * whole table (opticalimageprocessor_b200/sharding.py frames_index_shards).
 * oip_image_frames_hits: every occurrence of the trailer signature EB 90 E1 4D (ref aux_separator.h:82, :632 memmem) in
 * d_imdt[0, n_bytes), ascending, with the 172 bytes that start there (zero padded past the end); hits / trailers are HOST
 * arrays of capacity cap / cap * 172; *n_hits is set even when the capacity was too small (OIP_E_INVALID).
 * oip_image_frames_chain: host only, no context -- NextImageDataFrame + the gap rules of SeparateImageData over a
 * signature table whose offsets are relative to the whole stream of n_bytes (ref aux_separator.h:627-656, :287-320). */
int oip_image_frames_hits(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes, uint64_t *hits, uint8_t *trailers, int64_t cap,
                          int64_t *n_hits);
int oip_image_frames_chain(const uint64_t *hits, const uint8_t *trailers, int64_t n_hits, size_t n_bytes, const oip_frame_geom *geom,
                           oip_frame_entry *entries, int64_t cap, int64_t stats[4]);

/* aux copy + tile de-interleave + BE->LE swap for n_frames entries (zero fill for gap entries).
 * replaces WriteAuxData / WriteImageData / MergeSubImage / InflateSubImage(z_ratio==0) --
 * ref aux_separator.h:335-393.  Any of d_aux/d_pan/d_mss may be NULL to skip that product.
 * d_aux: n_frames * 192*tile_lines bytes; d_pan: n_frames*4*tile_lines lines of 8*tile_cols px;
 * d_mss: n_frames*tile_lines lines. */
int oip_unpack_frames(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes, const oip_frame_geom *geom,
                      const oip_frame_entry *entries, int64_t n_frames, uint8_t *d_aux,
                      uint16_t *d_pan, uint16_t *d_mss);

/* ---- stage 2: relative radiometric correction --------------------------------------------- */

/* dst = (uint16_t)(k[x]*src + b[x]) in fp64, truncation, in place.
 * replaces IMO::InplaceRRC(uint16_t* buff, int w, int h, const RRCParam*) -- ref imageop.h:129-138.
 * d_kb: w pairs {k,b} (RRCParam layout, ref imageop.h:26-29). */
int oip_rrc_u16(oip_ctx *ctx, uint16_t *d_img, int w, int64_t h, int64_t pitch_px, const double *d_kb);

/* host helper: parse an RRC CSV exactly like IMO::LoadRRCParamFile -- ref imageop.h:140-192 */
int oip_load_rrc_csv(const char *path, int expected_cols, double *h_kb);

/* ---- stage 3: stitch ----------------------------------------------------------------------- */

typedef enum {
    OIP_FMT_LE16 = 0,   /* u16 little-endian lines (.RAW, ref oipshared.h:27) */
    OIP_FMT_BE16 = 1,   /* u16 big-endian lines (byte order inside sub-images, ref aux_separator.h:387-392) */
    OIP_FMT_PACK12 = 2, /* extension: MSB-first 12-bit packed lines */
    OIP_FMT_PACK10 = 3, /* extension: MSB-first 10-bit packed lines */
    OIP_FMT_BE16_TILES = 4 /* image-frame sub-image layout straight from the IMDT stream */
} oip_sample_fmt;

/* rows of one CCD strip living in up to OIP_MAX_SEG device allocations (own shard, halo rows that
 * live on the neighbour GPUs and are read over NVLink, stale-section rows) */
#define OIP_MAX_SEG 4
typedef struct {
    const void *base;    /* first byte of row `row0` */
    int64_t row0;        /* global line index of the first row in this segment */
    int64_t n_rows;
    int64_t pitch_bytes; /* line formats: bytes between rows */
} oip_row_seg;

typedef struct {
    int fmt;             /* oip_sample_fmt */
    int n_seg;
    oip_row_seg seg[OIP_MAX_SEG];
    const double *d_kb;  /* w {k,b} pairs or NULL = no radiometric correction */
    int shifted;         /* 1: resample this CCD by (dX,dY) (CMOS-2.. in the reference); 0: copy (CMOS-1) */
    double dX, dY;       /* shift of this CCD relative to CCD 0 (used when shifted) */
    /* OIP_FMT_BE16_TILES only: seg[0].base = IMDT stream, d_tile_off[frame*40 + r*8+c] = byte offset
     * of each sub-image, -1 for a zero-filled frame; lines_per_frame = 4*tile_lines */
    const int64_t *d_tile_off;
    int tile_cols, tile_lines;
    /* OIP_FMT_BE16_TILES, optional: the same table in HOST memory (oip_image_frames_index returns its entries on the
     * host).  With it the planner can prove the sub-images 4-byte aligned and sends the CCD through the fast kernel,
     * which gathers its stage rows straight from the sub-images; without it the CCD runs on the generic kernel. */
    const int64_t *h_tile_off;
} oip_ccd_src;

typedef struct {
    int n_ccd;           /* 1..8 */
    int w;               /* pixels per line per CCD [12288] */
    int64_t total_rows;  /* lines of the whole strip (section geometry is global) */
    int64_t row0;        /* first output line produced by this call (multi-GPU shard) */
    int64_t n_rows;      /* output lines produced by this call */
    int fold_half;       /* f = fold_cols/2 (ref main.cpp:189) */
    int section_rows;    /* REMAP_SECTION_ROWS [30000] (ref imageop.h:20) */
    int row_guard;       /* REMAP_ROW_GUARD [32767] (ref imageop.h:19) */
    oip_ccd_src ccd[8];
    uint16_t *d_out;     /* n_rows x out_w, row `row0` first */
    int64_t out_pitch_px;
} oip_pan_desc;

/* Fused unpack -> RRC -> sub-pixel shift -> trimmed concat, one launch.
 * replaces, in one pass and without the .RRC.RAW / .PRESTT.RAW intermediates:
 *   IMO::DoRRC4RAW/InplaceRRC           ref imageop.h:194-228, :129-138
 *   Stitcher::PreStitch + IMO::SectionaryRemap + cv::remap(INTER_CUBIC,BORDER_CONSTANT)
 *                                       ref stitcher.h:83-139, imageop.h:230-275
 *   IMO::StitchBigRaw (RAW writer)      ref imageop.h:277-363
 * out_w = n_ccd*w - 2*(n_ccd-1)*fold_half.  total_rows <= row_guard: one section (the reference
 * throws, ref imageop.h:242-244). */
int oip_pan_pipeline(oip_ctx *ctx, const oip_pan_desc *desc);
int oip_pan_out_width(int n_ccd, int w, int fold_half);
/* synchronises and reports OIP_E_RANGE if a kernel needed a source row no segment supplied */
int oip_pan_check_error(oip_ctx *ctx);
/* the 32x4 bicubic weight table the kernels use (OpenCV interpolateCubic, A=-0.75) */
void oip_cubic_tab(float *tab128);
/* source rows [*first,*last) of CCD `ccd` that producing [row0,row0+n_rows) reads (halo planning) */
int oip_pan_rows_needed(const oip_pan_desc *desc, int ccd, int64_t *first, int64_t *last,
                        int64_t *stale_first, int64_t *stale_last);
/* the same as up to max_ranges disjoint, sorted [first,last) pairs in ranges[2*max_ranges] (ranges closer than 64 rows
 * are merged): what a host-buffer pipeline has to copy for one row block */
int oip_pan_row_ranges(const oip_pan_desc *desc, int ccd, int64_t *ranges, int max_ranges, int *n_ranges);

/* host-only planning diagnostic (no device work): how oip_pan_pipeline would split the output between its
 * regular-interior fast kernel and the exact generic kernel.  cover (n_rows x out_pitch_px bytes, zeroed by the
 * caller, may be NULL): += 1 per generic tile pixel, += 2 per fast tile pixel.
 * stats = {generic px, fast px, generic tiles, fast warp-tiles}. */
int oip_pan_plan_coverage(const oip_pan_desc *desc, int enable_fast, int fast_rows, uint8_t *cover, int64_t stats[4]);

/* stand-alone forms of the same kernel (same code path, one CCD / no shift) */
/* replaces Stitcher::PreStitch + IMO::SectionaryRemap -- ref stitcher.h:83-139, imageop.h:230-275 */
int oip_shift_cubic_u16(oip_ctx *ctx, const uint16_t *d_src, uint16_t *d_dst, int w, int64_t rows,
                        double dX, double dY, int section_rows, int row_guard);
/* replaces IMO::StitchBigRaw(left,right,out,pixelPerLine,foldColPixels) -- ref imageop.h:277-363 */
int oip_stitch_concat_u16(oip_ctx *ctx, const uint16_t *const *d_ccd, int n_ccd, int w, int64_t rows,
                          int fold_half, uint16_t *d_dst);

/* MSS: band split + per-band RRC + per-band polynomial cubic remap + 4-channel merge, sectioned.
 * replaces PreProcessor::LoadMSS (split), DoRRC4MSS, DoInterBandAlignment (both overloads) --
 * ref preproc.h:56-80, :202-222, :351-468.
 * d_mss: `lines` lines of 4*wb px (band b = columns [b*wb,(b+1)*wb)), pitch_px apart.
 * d_kb[b]: wb {k,b} pairs or NULL.  cX[b*2+i], cY[b*3+i] = mDeltaXcoeffs/mDeltaYcoeffs (ref :597-598).
 * d_out: (lines - line_offset - (keep_leading?0:overlap)) rows x wb x 4 interleaved u16
 * (cv::merge layout, ref :464); rows the reference never writes are left untouched.
 * *rows_written = processedLines. */
typedef struct {
    int fmt;              /* OIP_FMT_LE16 or OIP_FMT_BE16 */
    int wb;               /* PIXELS_PER_MSSBAND [3072] */
    int64_t lines;
    int64_t pitch_px;
    const double *d_kb[4];
    double cX[8];
    double cY[12];
    int lines_per_section;  /* IBPA_DEFAULT_BATCHLINES [20000] */
    int64_t line_offset;    /* [0] */
    int overlap;            /* IBPA_DEFAULT_LINEOVERLAP [520] */
    int keep_leading;       /* -k */
    int min_process_lines;  /* IBPA_MIN_PROCESSLINES [1500] */
    /* section shard: the sections of the reference loop are independent, so a rank may produce a subset.
     * sec_count == 0: all sections (then src_row0 must be 0).  Otherwise sections [sec_first, sec_first + sec_count) of
     * the `lines`-line strip; d_mss points at strip line src_row0 (the rank holds at least the lines of its sections)
     * and d_out row 0 is the first output row of section sec_first. */
    int sec_first, sec_count;
    int64_t src_row0;
} oip_mss_desc;
int oip_band_align_merge(oip_ctx *ctx, const void *d_mss, const oip_mss_desc *desc, uint16_t *d_out,
                         int64_t *rows_written);

/* host-only planning diagnostic (no device work): how oip_band_align_merge splits its output between the
 * regular-interior fast kernel and the exact generic kernel.  cover (rows_out x wb x 4 bytes, zeroed by the caller,
 * may be NULL): += 1 per generic sample, += 2 per fast sample.  stats = {generic samples, fast samples, generic tiles,
 * fast warp-tiles} (per band sample = one u16 of the interleaved raster). */
int oip_mss_plan_coverage(const oip_mss_desc *desc, int enable_fast, int tile_rows, uint8_t *cover, int64_t stats[4]);

/* replaces the geometry of IMO::StitchTiff / StitchTiffGDAL on CV_16UC4 data incl. the 1-based
 * band map -- ref imageop.h:416-421, :501-506, :529 */
int oip_stitch_concat_c4(oip_ctx *ctx, const uint16_t *const *d_img, int n_img, int w, int64_t rows,
                         int fold_half, const int *band_map, uint16_t *d_dst);

/* extension: unpack MSB-first packed 10/12-bit lines (or swap BE16) to u16 LE */
int oip_unpack_lines(oip_ctx *ctx, const void *d_in, int fmt, int w, int64_t rows, int64_t pitch_bytes,
                     uint16_t *d_out);

/* ---- inter-CMOS offset estimation (the caller that produces dX, dY) ---------- */
/* replaces cv::phaseCorrelate(src1, src2, noArray(), &response) as called at ref stitcher.h:180 on two u16 slices
 * (converted to float like the Mat1w -> Mat1f assignment at :175-176).  result = {dx, dy, response}.
 * Floating point: agrees with OpenCV within ~1e-3 px (different DFT), not bit for bit.  Even and odd optimal DFT sizes. */
int oip_phase_correlate_u16(oip_ctx *ctx, const uint16_t *d_a, int64_t pitch_a_px, const uint16_t *d_b, int64_t pitch_b_px,
                            int rows, int cols, double result[3]);

typedef struct oip_stt_config { /* ref oipshared.h:49-54, Stitcher ctor stitcher.h:53-58 */
    int32_t sections;          /* STT_DEF_SECTIONS 10 */
    int32_t lines_per_section; /* STT_DEF_SECLINES 16000 */
    int32_t overlap_cols;      /* STT_DEF_OVERLAPPX 200 */
    int32_t edge_cols;         /* STT_DEF_EDGECOLS 0 */
    double threshold;          /* STT_DEF_PHCTHRHLD 0.4 */
    double max_delta_y;        /* STT_DEF_MAXDELTAY 0.0 = no filter */
} oip_stt_config;
typedef struct oip_stt_section {
    int64_t line_offset;       /* first line of the section (global) */
    double dx, dy, response;
    int32_t valid;             /* 1 accepted, 0 rejected (threshold / max_delta_y), -1 rows not held by this shard */
    int32_t pad;
} oip_stt_section;
/* replaces Stitcher::CalcSttParameters -- ref stitcher.h:148-201.  d_pan1 / d_pan2 hold rows [row0, row0+rows_here) of
 * the two RRC-corrected strips (w px per line); sections wholly inside are correlated, the others are marked -1 (a
 * scanline-block shard passes its own rows; ranks add their sums[] = {sum dx, sum dy, sum response, n valid} with one
 * small all-reduce and divide, :197-199).  sections_out has cfg->sections entries. */
int oip_stt_parameters(oip_ctx *ctx, const uint16_t *d_pan1, const uint16_t *d_pan2, int w, int64_t total_lines,
                       int64_t row0, int64_t rows_here, int64_t pitch_px, const oip_stt_config *cfg,
                       oip_stt_section *sections_out, double sums[4]);

/* ---- inter-band shift estimation + polynomial fit (produces cX, cY of oip_mss_desc) ---- */
typedef struct oip_ibc_config {  /* ref oipshared.h:33-39 */
    int32_t slices;              /* IBCV_DEF_SLICES 10 (>= min_slices) */
    int32_t sections;            /* IBCV_DEF_SECTIONS 5 */
    double threshold;            /* IBCV_DEF_THRESHOLD 0.4 */
    int32_t correlation_lines;   /* CORRELATION_LINES 16000 (0 = default) */
    int32_t min_slices;          /* IBCV_MIN_SLICES 8 (0 = default) */
    int32_t min_count;           /* IBCV_MIN_COUNT 5 (0 = default) */
    int32_t pad;
} oip_ibc_config;
typedef struct oip_ibc_shift {   /* InterBandShift, ref preproc.h:23-28 */
    double dx, dy, rs;           /* dx, dy = NaN when rs < threshold (FilterInterBandShiftValues) */
    int32_t cx;                  /* centre column of the slice, PAN pixels */
    int32_t pad;
} oip_ibc_shift;
/* replaces PreProcessor::CalcInterBandCorrelation + FilterInterBandShiftValues + DoCorrelationPolynomialFitting --
 * ref preproc.h:224-347, :492-550.  d_pan = the (RRC'd) PAN strip, w px per line; d_mss = the MSS strip, each line the
 * 4 bands side by side (ref preproc.h:62-75).  Every PAN slice is correlated with the band slice upscaled x4 by
 * cv::resize INTER_CUBIC semantics; shifts[band * slices*sections + sec*slices + i]; cX[band*2 + k], cY[band*3 + k]
 * ascending coefficients as in mDeltaXcoeffs / mDeltaYcoeffs.  Fewer than min_count usable values in a band ->
 * OIP_E_RANGE with the reference's message.  Floating point: tolerance, not bits. */
int oip_inter_band_correlation(oip_ctx *ctx, const uint16_t *d_pan, int w, int64_t lines_pan, int64_t pan_pitch_px,
                               const uint16_t *d_mss, int64_t lines_mss, int64_t mss_pitch_px, const oip_ibc_config *cfg,
                               oip_ibc_shift *shifts, double cX[8], double cY[12]);

/* ---- raw downlink files -> stitched PAN raster in one call ------------------- */
/* one CCD's downlink: the AOS file image (ref aux_separator.h:224-245 mmaps it), device resident */
typedef struct {
    const uint8_t *d_file;
    size_t n_bytes;
    const double *d_kb;  /* 8*tile_cols {k,b} pairs or NULL */
    int shifted;         /* as in oip_ccd_src */
    double dX, dY;
} oip_downlink_src;
typedef struct {
    int n_ccd;           /* 1..8 */
    oip_frame_geom geom;
    int fold_half, section_rows, row_guard;   /* as in oip_pan_desc */
    oip_downlink_src ccd[8];
    uint16_t *d_out;     /* out_rows_cap x out_w, out_w = oip_pan_out_width(n_ccd, 8*tile_cols, fold_half) */
    int64_t out_pitch_px;
    int64_t out_rows_cap;
    uint8_t *d_aux[8];   /* optional per CCD: n_frames x 192*tile_lines bytes (.AUX, ref aux_separator.h:335-339) */
    uint16_t *d_mss[8];  /* optional per CCD: n_frames*tile_lines lines of 8*tile_cols px (.MSS.RAW, ref :341-364) */
} oip_downlink_desc;
typedef struct {
    int64_t aos[3];      /* valid, invalid, empty AOS frames (ref aux_separator.h:411-413) */
    int64_t imtr[9];     /* as oip_imtr_deframe */
    int64_t frames[4];   /* as oip_image_frames_index */
    int64_t imdt_bytes;
} oip_downlink_stats;
/* replaces, for the PAN product, the whole chain `auxsep` -> `prestitch` -> `stitch`:
 *   AuxSeparator::Separate (SeparateAosFile, DataTransFrameParser, SeparateImageData)  ref aux_separator.h:224-245, :256-590
 *   IMO::InplaceRRC, Stitcher::PreStitch + IMO::SectionaryRemap, IMO::StitchBigRaw       ref imageop.h:129-138, :230-363
 * The fused PAN kernel reads the sub-images where they lie in the IMDT stream; .PAN.RAW / .RRC.RAW / .PRESTT.RAW never
 * exist.  *rows_out = lines produced (frames x 4*tile_lines, the shortest CCD decides); stats[n_ccd] may be NULL.
 * Stage 1 of the CCDs runs side by side (one child context with its own stream and one host thread per CCD, joined
 * before the PAN launch on ctx's stream; the files are read after whatever ctx's stream held when the call was made).
 * A failure in one CCD is returned with "ccd <i>: " in front of its text. */
int oip_downlink_to_stitched(oip_ctx *ctx, const oip_downlink_desc *desc, int64_t *rows_out, oip_downlink_stats *stats);

/* ---- bench / test input (NOT a replaced reference function) --------------------------------- */
/* The reference ships no sample data; SURVEY 8(d) defines the synthetic strip DN(x,y) = 64 + ((37x mod 1500 + (y/8) mod 1200 +
 * (splitmix64(seed ^ (y*w + x)) & 0xFF)) mod 3968).  This fills rows [row0, row0+rows) of that strip on the device
 * (bit-identical to opticalimageprocessor_b200/synth.py:strip_dn), little- or big-endian samples: how bench.py makes its
 * 51.5 GB C4 input without a host generator. */
int oip_synth_strip_dn(oip_ctx *ctx, uint16_t *d_out, int w, int64_t rows, int64_t row0, int64_t pitch_px,
                       uint64_t seed, int big_endian);

/* ---- whole-stage host-buffer entry points (what the CLI and bench.py "e2e" call) ------------ */
/* host in / host out, copies on side streams overlapped with the kernels in row blocks */
int oip_pan_pipeline_host(oip_ctx *ctx, const oip_pan_desc *desc_host_ptrs);

#ifdef __cplusplus
}
#endif
#endif /* OIP_B200_H */
