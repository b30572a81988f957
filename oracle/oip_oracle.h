/*
 * oip_oracle.h -- CPU restatement of the OpticalImageProcessor pre-processing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (opticalimageprocessor_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - the reference ships no tests and no golden vectors;
 *   - CRC: pinned against the reference's own vendored CRC.h compiled stand-alone
 *     (oracle/_ref/libref_crc.so) and its check value 0x29B1 (CRC.h:1519);
 *   - resampling: pinned against cv2.remap 4.13.0 (the library call the reference makes at
 *     imageop.h:258 / preproc.h:453) via tests/golden fixtures;
 *   - frame handling / sectioning control flow: pinned against the reference's own headers
 *     compiled with stub third-party headers (oracle/_ref/libref_oip.so) where that builds.
 *
 * "ref" in comments = /root/reference/OpticalImageProcessor/<file>:<line>.
 */
#ifndef OIP_ORACLE_H
#define OIP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- stage 1: frame handling (aux_separator.h, CRC.h) ---------------------------------- */

/* CRC-16/CCITT-FALSE, bit-by-bit. ref CRC.h:806-834 (MSB-first branch), :707-720, :1522-1526 */
uint16_t oipo_crc16(const uint8_t *data, size_t n);

/* ValidateAosFrame. ref aux_separator.h:658-690. returns -1 invalid, 0 empty, 1 valid */
int oipo_aos_validate(const uint8_t *frame, uint32_t *vcid, uint32_t *seq, uint32_t *inj,
                      uint32_t *crc_in_frame);

/* SeparateAosFile chained scan. ref aux_separator.h:395-467, :622-625.
 * payload_off[i] = byte offset (into buf) of the 880-byte payload of the i-th VALID frame.
 * counters = {valid, invalid, empty}.  A sync hit counts only if off+1024 <= n (SURVEY C-5).
 * returns number of valid frames (may exceed cap; only cap entries are written). */
int64_t oipo_aos_scan(const uint8_t *buf, size_t n, uint64_t *payload_off, size_t cap,
                      int64_t counters[3]);

/* byte-range shard of the scan: candidates starting in [start, own_end); see oip_oracle.c */
int64_t oipo_aos_scan_range(const uint8_t *buf, size_t n, size_t start, size_t own_end, uint64_t *payload_off, size_t cap,
                            int64_t counters[3], uint64_t *next_pos);

/* DataTransFrameParser + ValidateImtrFrame. ref aux_separator.h:469-590.
 * Cuts the concatenated 880-byte payloads at a fixed 882-byte cadence from stream byte 0,
 * validates, appends the 866-byte body of each valid frame to imdt.
 * stats = {frames_cut, frames_valid, bad_sig, bad_endsig, bad_type, bad_crc, seq_gaps,
 *          first_chid, restarts}  ("restarts": the IMDT file is (re)created whenever the
 * previously accepted frame had seq 0, ref :513-528 -- output holds bytes since the last one).
 * returns bytes written to imdt. */
int64_t oipo_imtr_deframe(const uint8_t *buf, const uint64_t *payload_off, int64_t n_payload,
                          uint8_t *imdt, size_t cap, int64_t stats[9]);

/* geometry of an image frame (ref aux_separator.h:80-118); reference values in brackets */
typedef struct {
    int tile_cols;   /* IMGSIG_IMBASE_COLS  [1536] */
    int tile_lines;  /* IMGSIG_IMBASE_LINES [256]  */
    /* fixed by the 40-entry trailer: 8 horizontal parts, 4 PAN + 1 MSS vertical parts */
} oipo_frame_geom;

/* SeparateImageData + NextImageDataFrame + WriteImageData. ref aux_separator.h:256-393, :627-656.
 * Outputs are appended: aux (48*4*tile_lines bytes/frame), pan (4*tile_lines lines of
 * 8*tile_cols u16 LE), mss (tile_lines lines).  Gap frames are zero-filled (ref :302-311).
 * Only uncompressed frames (z_ratio==0) are supported; a compressed frame returns -2.
 * A tile read that would leave the buffer returns -3 (the reference would read out of bounds).
 * stats = {frames_found, frames_emitted (incl. zero-filled), frames_incomplete, last_seq}.
 * returns number of emitted frames (incl. zero-filled) or <0. Pass NULL outputs to count only. */
int64_t oipo_image_frames(const uint8_t *imdt, size_t n, const oipo_frame_geom *g,
                          uint8_t *aux, uint16_t *pan, uint16_t *mss, int64_t cap_frames,
                          int64_t stats[4]);

/* ---- stage 2: relative radiometric correction (imageop.h) ------------------------------- */

/* InplaceRRC. ref imageop.h:129-138.  kb = {k0,b0,k1,b1,...} (RRCParam layout, ref :26-29) */
void oipo_rrc_u16(uint16_t *buf, int w, int64_t h, const double *kb);

/* LoadRRCParamFile. ref imageop.h:140-192. returns 0 or <0 */
int oipo_load_rrc_csv(const char *path, int expected, double *kb);

/* LoadMSS band split. ref preproc.h:56-80 */
void oipo_mss_split(const uint16_t *mixed, int64_t lines, int line_px, uint16_t *planes[4]);

/* ---- stage 3: stitch (stitcher.h, imageop.h, preproc.h, OpenCV remap semantics) ---------- */

/* cv::remap(INTER_CUBIC, BORDER_CONSTANT 0) on CV_16UC1 with float maps, as called at
 * imageop.h:258 / preproc.h:453.  OpenCV 4.x semantics restated.
 * mapx/mapy: dh x dw floats. */
void oipo_remap_cubic_u16(const uint16_t *src, int sw, int sh, int64_t sstep_px, uint16_t *dst,
                          int dw, int dh, const float *mapx, const float *mapy);

/* the 32x4 1-D cubic weight table OpenCV builds (interpolateCubic, A=-0.75) */
void oipo_cubic_tab(float tab[32 * 4]);

/* Stitcher::PreStitch + IMO::SectionaryRemap, literal (30000-row reused buffer, ucut/bcut,
 * stale rows of a partial last section).  ref stitcher.h:83-139, imageop.h:230-275.
 * section_rows [30000], row_guard [32767].  total_rows <= row_guard: the reference throws
 * (imageop.h:242-244); here: one cv::remap over the whole image (documented extension).
 * returns rows written (== total_rows) or <0 */
int64_t oipo_prestitch_shift(const uint16_t *src, int w, int64_t total_rows, double dX, double dY,
                             int section_rows, int row_guard, uint16_t *dst);

/* N-CCD generalisation of StitchBigRaw: ccd0 keeps [0,W-f), middle [f,W-f), last [f,W).
 * ref imageop.h:291-295, :340-355; main.cpp:189. n_ccd==2 is the reference. */
void oipo_stitch_concat_u16(const uint16_t *const *ccd, int n_ccd, int w, int64_t h, int fold_half,
                            uint16_t *dst);

/* whole PAN path = RRC each CCD, shift CCD i>=1 by (dX[i],dY[i]), concat. in: LE u16 lines */
int64_t oipo_pan_pipeline(const uint16_t *const *ccd, int n_ccd, int w, int64_t h,
                          const double *const *kb, const double *dX, const double *dY,
                          int fold_half, int section_rows, int row_guard, uint16_t *dst);

/* PreProcessor::DoInterBandAlignment (both overloads), literal.  ref preproc.h:351-468.
 * planes: 4 band planes lines x wb.  cX[b][2], cY[b][3].  out: (lines-line_offset-(keep?0:overlap))
 * rows x wb x 4 interleaved; rows never written stay 0 (the reference leaves them
 * uninitialised, SURVEY C-4).  returns rows written (processedLines) or <0 on the reference's
 * argument errors. */
void oipo_set_min_process_lines(int v); /* IBPA_MIN_PROCESSLINES [1500], oipshared.h:46 */
int64_t oipo_band_align(const uint16_t *const planes[4], int64_t lines, int wb,
                        const double cX[4][2], const double cY[4][3], int lines_per_section,
                        int64_t line_offset, int overlap, int keep_leading, uint16_t *out);

/* StitchTiff geometry on 4-channel pixels with optional 1-based band map. ref imageop.h:416-421,
 * :501-506, :529.  n_img generalises 2. */
void oipo_stitch_concat_c4(const uint16_t *const *img, int n_img, int w, int64_t h, int fold_half,
                           const int *band_map /* NULL or 4 ints 1-based */, uint16_t *dst);

/* ---- packed-sample extension (not in the reference) --------------------- */
/* MSB-first big-endian bitstream of `bits`-wide samples, each line starts byte aligned. */
void oipo_unpack_bits(const uint8_t *in, int bits, int w, int64_t h, int64_t pitch_bytes,
                      uint16_t *out);
void oipo_swap16(const uint16_t *in, int64_t n, uint16_t *out); /* ref aux_separator.h:387-392 */

#ifdef __cplusplus
}
#endif
#endif
