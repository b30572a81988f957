// mss_plan.cuh -- tile descriptions shared by the MSS kernels (mss.cu generic, mss_fast.cu regular-interior fast path)
#pragma once
#include <vector>

#include "tma_warp.cuh"

namespace oip {
namespace mss {
// generic tile: any rectangle of one band inside one section
struct Tile {
    int32_t band, x_begin, x_end;
    int32_t rows;       // section height = rows of the cv::Mat handed to cv::remap (ref preproc.h:453)
    int32_t y0, n_rows; // section-local output rows [y0, y0+n_rows)
    int64_t sec_off;    // first source line of the section (rowOffset)
    int64_t dst_row0;   // output raster row of y0
};
} // namespace mss

namespace mssfast {
constexpr int WARPS = 4;
struct FTile {
    int32_t band;      // -1: padding entry
    int32_t x_begin;   // first output column (band-relative)
    int32_t nh;        // left half = columns [x_begin, x_begin+nh), right half = [x_begin+nh, x_begin+nh+n_right)
    int32_t n_right;   // <= nh
    int32_t ix0;       // band-relative source column of x_begin's first tap
    int32_t ya;        // section-local row of output row 0 (the device derives every column's phase from it)
    int32_t n_rows;
    int32_t pad;       // 1: the strip touches the band border (EDGE variant of the kernel)
    int64_t src_row0;  // line of the MSS buffer that holds the first tap row of output row 0
    int64_t out_off;   // element offset of (output row 0, x_begin, band) in the interleaved raster
};
struct Params {
    CUtensorMap tmap;
    const double *kb[4];
    double cX[8], cY[12];
    const FTile *tiles;
    uint16_t *out;
    const float *tab; // 32x4 cubic weights
    uint64_t nz;      // run-time (-0.0,-0.0) addend of the packed products: a uniform-register operand (see pan_fast.cuh)
    int32_t wb, swap, n_stage, pad;
};
// one section of the reference's loop (ref preproc.h:379-408)
struct Section {
    int64_t sec_off;  // first source line of the section Mat
    int rows;         // its height
    int y0;           // first kept row
    int64_t dst_row0; // output raster row of y0
};
void plan(const oip_mss_desc *d, const std::vector<Section> &secs, bool fast, int tile_rows, std::vector<mss::Tile> &tiles,
          std::vector<FTile> &ftiles);
int launch(oip_ctx *ctx, const Params &P, int64_t n_ctas);
int encode(CUtensorMap *tm, const void *base, int line_px, int64_t lines, int64_t pitch_bytes);
} // namespace mssfast
} // namespace oip
