// pan_fast.cuh -- warp-tile description shared by the planner (pan_pipeline.cu) and the fast kernel (pan_fast.cu)
#pragma once
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "oip_common.cuh"

namespace oip {
namespace panfast {

constexpr int WARPS = 4;        // independent warp workers per CTA (one warp-tile each)
constexpr int RC = 4;           // source rows per TMA stage
constexpr int BOX_W = 136;      // 32-bit elements (sample pairs) per box row: one box of 272 samples x RC rows per stage
constexpr int HALF_MAX = 124;   // REMAP: output columns per half (31 lanes x 4); window = half + 3 (+1 slack) columns
constexpr int COPY_MAX = 256;   // COPY: output columns per warp-tile (32 lanes x 8)
constexpr int MAX_MAPS = 8 * OIP_MAX_SEG;

enum { FT_NONE = -1, FT_COPY = 0, FT_REMAP = 1, FT_EDGE = 2 };
constexpr int EDGE_MAX = 29;    // EDGE: output columns per warp-tile (one per lane; 3 more lanes supply window columns)

// one warp's work: a column strip of one CCD over n_rows output rows whose source rows sit in ONE row
// segment, whose 4x4 footprints are all inside the section's fresh rows / the CCD's columns, and whose
// fixed-point map advances by exactly 32 per output pixel and per output row ("regular")
struct FastTile {
    int32_t kind;     // FT_*
    int32_t ccd;
    int32_t tmap;     // line formats: tensor map index = ccd * OIP_MAX_SEG + segment;  tiled CCDs: sub-image column floor(box origin / tile_cols)
    int32_t x_begin;  // first CCD column produced
    int32_t half;     // EDGE: columns of the sliver (<= EDGE_MAX), any alignment, image-border columns allowed;  REMAP: columns per half (multiple of 4, <= HALF_MAX), n_cols = 2*half; COPY: n_cols (any, <= COPY_MAX)
    int32_t src_x0;   // REMAP / EDGE: source column of the first tap of x_begin (EDGE: may be < 0);  COPY: x_begin
    int32_t src_y0;   // row inside the segment of the first source row (first tap row of output row 0)
    int32_t n_rows;   // output rows
    int32_t fx, fy;   // sub-pixel phase (1/32) of the whole tile
    int64_t out_off;  // element offset in the output raster of (output row 0, x_begin)
};

// exact n / d for n < 2^32 by one multiply-high (Granlund-Montgomery round-up form): q = (t + ((n - t) >> s1)) >> s2, t = mulhi(m, n)
struct FastDiv { uint32_t m, s1, s2, d; };
inline FastDiv fast_div_make(uint32_t d)
{
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    FastDiv f;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l < 1 ? 0 : l - 1;
    f.d = d;
    return f;
}

struct FastCcd {
    const double *kb; // {k,b} per detector or null
    int32_t swap;     // 1: samples are big-endian
    int32_t tiled;    // 1: OIP_FMT_BE16_TILES -- rows are gathered from the sub-images of the IMDT stream (ref aux_separator.h:341-372)
    int32_t pbits;    // 0, or 12 / 10: MSB-first packed samples (OIP_FMT_PACK12 / PACK10): the stage is unpacked in shared memory
    int32_t pad2;
    // tiled sources only
    const uint8_t *tile_base;  // IMDT stream (4-byte aligned)
    const int64_t *tile_off;   // [frame * 40 + r * 8 + c] byte offset of each sub-image (multiple of 4), -1 = zero-filled frame
    int32_t tile_cols, tile_lines, n_frames, pad;
    FastDiv div_lpf, div_tl;   // / (4 * tile_lines), / tile_lines
};

struct FastParams {
    CUtensorMap tmap[MAX_MAPS];
    FastCcd ccd[8];
    const FastTile *tiles;
    uint16_t *out;
    int64_t out_pitch;
    const float *tab; // 32x4 cubic weights, then the run-time -0.0 addend of the packed products (mul2())
    int32_t w;
    int32_t n_stage; // TMA stages per warp (2..8)
};

// host side (pan_fast.cu)
int fast_encode_tmap(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes);
int fast_encode_tmap_packed(CUtensorMap *tm, const void *base, int w, int bits, int64_t n_rows, int64_t pitch_bytes);
constexpr int PBOX12 = 116, PBOX10 = 108; // 32-bit elements per box row of a packed stage: 16 + 288 samples of 12 bits, 48 + 288 of 10 bits
// source-format class of a CCD: one pan_fast_kernel instantiation (and translation unit, pan_fast_c<N>.cu) each
inline int fast_class(int fmt) { return fmt == OIP_FMT_BE16_TILES ? 1 : (fmt == OIP_FMT_PACK12 || fmt == OIP_FMT_PACK10) ? 2 : 0; }
template <int CLS> int fast_launch_cls(oip_ctx *ctx, const FastParams &P, int64_t tile0, int64_t n_ctas);
template <> int fast_launch_cls<0>(oip_ctx *ctx, const FastParams &P, int64_t tile0, int64_t n_ctas);
template <> int fast_launch_cls<1>(oip_ctx *ctx, const FastParams &P, int64_t tile0, int64_t n_ctas);
template <> int fast_launch_cls<2>(oip_ctx *ctx, const FastParams &P, int64_t tile0, int64_t n_ctas);
int fast_launch(oip_ctx *ctx, const FastParams &P, const int64_t *cls_ctas); // the classes follow each other in P.tiles

} // namespace panfast
} // namespace oip
