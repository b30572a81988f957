// mss.cu -- MSS: band split + per-band RRC + per-band polynomial bicubic remap + 4-channel merge.
// replaces PreProcessor::LoadMSS (split), DoRRC4MSS, DoInterBandAlignment -- ref preproc.h:56-80,
// :202-222, :351-468.  The band split is free (a band is a column range of the mixed line), RRC is
// applied while a source row is staged in shared memory, and the remapped value is stored straight
// into its channel of the interleaved CV_16UC4 raster (cv::merge layout, ref preproc.h:464).
#include <algorithm>

#include "mss_plan.cuh"

namespace oip {
namespace mss {

// tile limits (the planner in mss_fast.cu cuts generic tiles to them): 240 output columns x 256 output rows
constexpr int SWC = 256;  // staged source columns
constexpr int RC = 32;    // output rows per chunk
constexpr int RING = 48;  // float ring rows: RC + 3 taps + spread of the per-column row offsets
constexpr int NT = 256;

struct Params {
    const uint8_t *base;
    int64_t pitch_bytes;
    const double *kb[4];
    double cX[8], cY[12];
    const Tile *tiles;
    uint16_t *out;
    const float *tab;
    int *err;
    int32_t fmt, wb, vec_ok;
};

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

// block-wide min / max of an int (all threads participate)
__device__ __forceinline__ void block_minmax(int v_min, int v_max, int *s_red, int &out_min, int &out_max)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        v_min = min(v_min, __shfl_xor_sync(0xffffffffu, v_min, o));
        v_max = max(v_max, __shfl_xor_sync(0xffffffffu, v_max, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        s_red[wid] = v_min;
        s_red[8 + wid] = v_max;
    }
    __syncthreads();
    int a = s_red[0], b = s_red[8];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) {
        a = min(a, s_red[i]);
        b = max(b, s_red[8 + i]);
    }
    out_min = a;
    out_max = b;
}

__device__ __forceinline__ float dot4(const float (&v)[4], const float (&w)[4])
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], w[0]), __fmul_rn(v[1], w[1])), __fmul_rn(v[2], w[2])),
                     __fmul_rn(v[3], w[3]));
}
__device__ __forceinline__ float acc4(float s, const float (&v)[4], const float (&w)[4])
{
    s = __fadd_rn(s, __fmul_rn(v[0], w[0]));
    s = __fadd_rn(s, __fmul_rn(v[1], w[1]));
    s = __fadd_rn(s, __fmul_rn(v[2], w[2]));
    s = __fadd_rn(s, __fmul_rn(v[3], w[3]));
    return s;
}

__global__ void __launch_bounds__(NT, 2) band_align_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    float *ring = reinterpret_cast<float *>(smem); // RING x SWC
    __shared__ float s_tab[128];
    __shared__ int s_red[16];

    const int tid = threadIdx.x;
    const Tile T = P.tiles[blockIdx.x];
    const int b = T.band, wb = P.wb;
    const int n_cols = T.x_end - T.x_begin;
    const double cX0 = P.cX[2 * b], cX1 = P.cX[2 * b + 1];
    const double cY0 = P.cY[3 * b], cY1 = P.cY[3 * b + 1], cY2 = P.cY[3 * b + 2];
    if (tid < 128) s_tab[tid] = P.tab[tid];

    // ---- per-column map (ref preproc.h:443-449): fp64, left to right, then float
    const bool active = tid < n_cols;
    const int x = T.x_begin + (active ? tid : 0);
    const int xx = x * 4;                                                       // :446
    const double dxx = (double)xx;
    const float mapx = __double2float_rn(__ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(cX1, dxx), cX0), dxx), 4.0)); // :447
    const double Ay = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(cY2, dxx), dxx), __dmul_rn(cY1, dxx)), cY0);     // :448
    const int sx = __float2int_rn(__fmul_rn(mapx, 32.0f));
    const int ix = sat_short(sx >> 5) - 1, fx = sx & 31;
    auto sy_of = [&](int y) {
        const float m = __double2float_rn(__ddiv_rn(__dadd_rn(Ay, (double)((int64_t)y * 4)), 4.0));
        return __float2int_rn(__fmul_rn(m, 32.0f));
    };
    const int d_col = (sat_short(sy_of(T.y0) >> 5) - 1) - T.y0; // row offset of this column's first tap

    int ix_min, ix_max, d_min, d_max;
    block_minmax(active ? ix : INT_MAX, active ? ix : INT_MIN, s_red, ix_min, ix_max);
    block_minmax(active ? d_col : INT_MAX, active ? d_col : INT_MIN, s_red, d_min, d_max);
    d_min -= 1; // float rounding of the map can move a tap row by one over the tile's rows
    d_max += 1;
    const int c_lo = (ix_min >= 0 ? ix_min : ix_min - 7) / 8 * 8;
    if (ix_max + 3 >= c_lo + SWC || RC + 3 + (d_max - d_min) > RING) {
        // The polynomial is too steep for the staged window (the tap rows of the tile's columns spread over more than
        // RING rows, or its source columns over more than SWC): every thread resamples its column straight from global
        // memory, tap by tap -- slow, exact, and what the reference computes (round 1 returned OIP_E_UNSUPPORTED here).
        if (active) {
            const bool be_d = P.fmt == OIP_FMT_BE16;
            const uint8_t *bb = P.base + ((int64_t)b * wb) * 2;
            const double *kbd = P.kb[b];
            auto tap = [&](int t, int c) -> float {
                if (t < 0 || t >= T.rows || c < 0 || c >= wb) return 0.f; // constant border of the section Mat (SURVEY C-1)
                uint32_t a = *reinterpret_cast<const uint16_t *>(bb + (T.sec_off + t) * P.pitch_bytes + 2 * (int64_t)c);
                if (be_d) a = ((a & 0xFF) << 8) | (a >> 8);
                if (kbd) a = rrc_px(a, kbd[2 * c], kbd[2 * c + 1]);
                return u16_to_f32(a);
            };
            const bool col_int_d = (unsigned)ix < (unsigned)max(wb - 3, 0);
            uint16_t *o = P.out + ((T.dst_row0 * wb + x) * 4 + b);
            for (int i = 0; i < T.n_rows; ++i, o += (int64_t)wb * 4) {
                const int sy = sy_of(T.y0 + i);
                const int iy = sat_short(sy >> 5) - 1, fy = sy & 31;
                float wgt[4][4], v[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        wgt[r][c] = __fmul_rn(s_tab[4 * fy + r], s_tab[4 * fx + c]);
                        v[r][c] = tap(iy + r, ix + c);
                    }
                float sum;
                if ((unsigned)iy < (unsigned)max(T.rows - 3, 0) && col_int_d) {
                    sum = dot4(v[0], wgt[0]);
                    sum = __fadd_rn(sum, dot4(v[1], wgt[1]));
                    sum = __fadd_rn(sum, dot4(v[2], wgt[2]));
                    sum = __fadd_rn(sum, dot4(v[3], wgt[3]));
                } else {
                    sum = 0.f;
                    sum = acc4(sum, v[0], wgt[0]);
                    sum = acc4(sum, v[1], wgt[1]);
                    sum = acc4(sum, v[2], wgt[2]);
                    sum = acc4(sum, v[3], wgt[3]);
                }
                *o = (uint16_t)max(0, min(65535, __float2int_rn(sum)));
            }
        }
        return;
    }
    const bool col_int = (unsigned)ix < (unsigned)max(wb - 3, 0);
    const int cx = ix - c_lo;
    float wx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wx[i] = s_tab[4 * fx + i];

    // ---- convert-phase mapping (column pair, row parity) and its RRC coefficients
    const int p = tid & 127, par = tid >> 7;
    const int c0 = c_lo + 2 * p;
    const bool v0ok = c0 >= 0 && c0 < wb, v1ok = c0 + 1 >= 0 && c0 + 1 < wb;
    const double *kb = P.kb[b];
    double k0 = 1.0, b0 = 0.0, k1 = 1.0, b1 = 0.0;
    if (kb) {
        if (v0ok) { k0 = kb[2 * c0]; b0 = kb[2 * c0 + 1]; }
        if (v1ok) { k1 = kb[2 * c0 + 2]; b1 = kb[2 * c0 + 3]; }
    }
    const bool be = P.fmt == OIP_FMT_BE16;
    const uint8_t *band_base = P.base + ((int64_t)b * wb) * 2;

    const int n_chunks = (T.n_rows + RC - 1) / RC;
    const int t_base = T.y0 + d_min;
    int loaded_hi = t_base - 1; // last buffer row already in the ring
    uint16_t *out_px = P.out + ((T.dst_row0 * wb + x) * 4 + b);

    for (int k = 0; k < n_chunks; ++k) {
        const int ya = T.y0 + k * RC;
        const int nr = min(RC, T.n_rows - k * RC);
        const int t_hi = ya + nr - 1 + d_max + 3;
        // ---- stage + RRC rows (loaded_hi, t_hi]
        for (int t = loaded_hi + 1 + par; t <= t_hi; t += 2) {
            float2 o = make_float2(0.f, 0.f);
            if (t >= 0 && t < T.rows) { // rows outside the section Mat are constant border (SURVEY C-1)
                const uint8_t *row = band_base + (T.sec_off + t) * P.pitch_bytes;
                uint32_t a = 0, c = 0;
                if (P.vec_ok && v0ok && v1ok) {
                    uint32_t raw = *reinterpret_cast<const uint32_t *>(row + 2 * (int64_t)c0);
                    if (be) raw = bswap16x2(raw);
                    a = raw & 0xFFFFu;
                    c = raw >> 16;
                } else {
                    if (v0ok) { a = *reinterpret_cast<const uint16_t *>(row + 2 * (int64_t)c0); if (be) a = ((a & 0xFF) << 8) | (a >> 8); }
                    if (v1ok) { c = *reinterpret_cast<const uint16_t *>(row + 2 * (int64_t)c0 + 2); if (be) c = ((c & 0xFF) << 8) | (c >> 8); }
                }
                if (kb) {
                    a = rrc_px(a, k0, b0);
                    c = rrc_px(c, k1, b1);
                }
                o.x = v0ok ? u16_to_f32(a) : 0.f;
                o.y = v1ok ? u16_to_f32(c) : 0.f;
            }
            const int slot = (t - t_base) % RING;
            reinterpret_cast<float2 *>(ring)[(size_t)slot * (SWC / 2) + p] = o;
        }
        loaded_hi = t_hi;
        __syncthreads();

        // ---- resample
        if (active) {
            float wgt[4][4], v[4][4];
            int prev_iy = INT_MIN, prev_fy = -1;
            uint16_t *o = out_px + (int64_t)k * RC * wb * 4;
            for (int i = 0; i < nr; ++i, o += (int64_t)wb * 4) {
                const int sy = sy_of(ya + i);
                const int iy = sat_short(sy >> 5) - 1, fy = sy & 31;
                if (fy != prev_fy) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) wgt[r][c] = __fmul_rn(s_tab[4 * fy + r], wx[c]);
                    prev_fy = fy;
                }
                auto load_row = [&](int t, float(&dst)[4]) {
                    const bool in = t >= t_base && t <= t_hi && t > t_hi - RING;
                    const float *q = ring + (size_t)((t - t_base) % RING) * SWC + cx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[c] = in ? q[c] : 0.f;
                };
                if (iy == prev_iy + 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { v[0][c] = v[1][c]; v[1][c] = v[2][c]; v[2][c] = v[3][c]; }
                    load_row(iy + 3, v[3]);
                } else if (iy != prev_iy) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) load_row(iy + r, v[r]);
                }
                prev_iy = iy;
                const bool row_int = (unsigned)iy < (unsigned)max(T.rows - 3, 0);
                float s;
                if (row_int && col_int) {
                    s = dot4(v[0], wgt[0]);
                    s = __fadd_rn(s, dot4(v[1], wgt[1]));
                    s = __fadd_rn(s, dot4(v[2], wgt[2]));
                    s = __fadd_rn(s, dot4(v[3], wgt[3]));
                } else {
                    s = 0.f;
                    s = acc4(s, v[0], wgt[0]);
                    s = acc4(s, v[1], wgt[1]);
                    s = acc4(s, v[2], wgt[2]);
                    s = acc4(s, v[3], wgt[3]);
                }
                const int r = __float2int_rn(s);
                *o = (uint16_t)max(0, min(65535, r));
            }
        }
        __syncthreads();
    }
}

} // namespace mss
} // namespace oip

using namespace oip;

extern "C" void oip_cubic_tab(float *tab128);

// the reference's section loop (ref preproc.h:379-408); returns processedLines
static int64_t build_sections(const oip_mss_desc *d, std::vector<mssfast::Section> &secs)
{
    const int lps = d->lines_per_section, overlap = d->overlap;
    const int min_lines = d->min_process_lines > 0 ? d->min_process_lines : 1500;
    uint64_t offset = (uint64_t)d->line_offset;
    int64_t processed = 0;
    for (int i = 0;; ++i) {
        const uint64_t rem = (uint64_t)d->lines - offset;
        const uint64_t n = std::min<uint64_t>(rem, (uint64_t)lps);                    // :380
        if ((uint64_t)d->lines < offset || n < (uint64_t)min_lines) break;             // :381
        const int y0 = (i == 0 && d->keep_leading) ? 0 : overlap;                      // :392-402
        secs.push_back({(int64_t)offset, (int)n, y0, processed});
        processed += (int64_t)n - y0;                                                   // :396,405
        offset += (uint64_t)(lps - overlap);                                            // :407
    }
    if (d->sec_count > 0) { // section shard: keep [sec_first, sec_first + sec_count), relative to the shard's buffers
        std::vector<mssfast::Section> keep;
        int64_t out0 = 0, rows = 0;
        for (int i = d->sec_first; i < (int)secs.size() && i < d->sec_first + d->sec_count; ++i) {
            if (keep.empty()) out0 = secs[i].dst_row0;
            keep.push_back({secs[i].sec_off - d->src_row0, secs[i].rows, secs[i].y0, secs[i].dst_row0 - out0});
            rows += secs[i].rows - secs[i].y0;
        }
        secs.swap(keep);
        return rows;
    }
    return processed;
}

/* host-only diagnostic: how oip_band_align_merge splits its output between the fast and the generic kernel */
extern "C" int oip_mss_plan_coverage(const oip_mss_desc *d, int enable_fast, int tile_rows, uint8_t *cover, int64_t stats[4])
{
    if (!d || d->wb < 4 || d->lines < 0 || d->lines_per_section < 8 || d->overlap < 0 || d->lines_per_section <= d->overlap)
        return fail(OIP_E_INVALID, "oip_mss_plan_coverage: bad descriptor");
    std::vector<mssfast::Section> secs;
    build_sections(d, secs);
    std::vector<mss::Tile> tiles;
    std::vector<mssfast::FTile> ftiles;
    mssfast::plan(d, secs, enable_fast != 0, tile_rows, tiles, ftiles);
    int64_t g = 0, f = 0, nf = 0;
    const int64_t wb = d->wb;
    for (const mss::Tile &t : tiles) {
        g += (int64_t)(t.x_end - t.x_begin) * t.n_rows;
        if (cover)
            for (int r = 0; r < t.n_rows; ++r)
                for (int x = t.x_begin; x < t.x_end; ++x) cover[((t.dst_row0 + r) * wb + x) * 4 + t.band] += 1;
    }
    for (const mssfast::FTile &t : ftiles) {
        if (t.band < 0) continue;
        ++nf;
        f += (int64_t)(t.nh + t.n_right) * t.n_rows;
        if (cover)
            for (int r = 0; r < t.n_rows; ++r)
                for (int x = 0; x < t.nh + t.n_right; ++x) cover[t.out_off + (int64_t)r * wb * 4 + 4 * x] += 2;
    }
    if (stats) { stats[0] = g; stats[1] = f; stats[2] = (int64_t)tiles.size(); stats[3] = nf; }
    return OIP_OK;
}

extern "C" int oip_band_align_merge(oip_ctx *ctx, const void *d_mss, const oip_mss_desc *d, uint16_t *d_out,
                                    int64_t *rows_written)
{
    OIP_CHECK_CTX(ctx);
    if (rows_written) *rows_written = 0;
    if (!d || !d_mss || !d_out) return fail(OIP_E_INVALID, "oip_band_align_merge: null argument");
    if (d->fmt != OIP_FMT_LE16 && d->fmt != OIP_FMT_BE16) return fail(OIP_E_INVALID, "oip_band_align_merge: format %d", d->fmt);
    if (d->wb < 4 || d->wb > 32760 || d->lines < 0 || d->pitch_px < 4 * (int64_t)d->wb)
        return fail(OIP_E_INVALID, "oip_band_align_merge: bad geometry");
    const int lps = d->lines_per_section, overlap = d->overlap;
    const int min_lines = d->min_process_lines > 0 ? d->min_process_lines : 1500;
    // the reference's argument checks, same order and wording (ref preproc.h:355-367)
    if (overlap > 3000) return fail(OIP_E_INVALID, "Overlap value %d exceeds maximum allowed value(%d)", overlap, 3000);
    if (lps > 32767) return fail(OIP_E_INVALID, "Row number exceeds OpenCV allowed value");
    if (lps < overlap * 2) return fail(OIP_E_INVALID, "Lines per section too small or section overlapped lines too large");
    if (d->lines - d->line_offset < min_lines) return fail(OIP_E_INVALID, "Too few image lines left to process");
    if (overlap < 0 || d->line_offset < 0) return fail(OIP_E_INVALID, "negative overlap / line offset");
    if (d->sec_count < 0 || d->sec_first < 0 || d->src_row0 < 0 || (d->sec_count == 0 && d->src_row0 != 0))
        return fail(OIP_E_INVALID, "oip_band_align_merge: bad section shard (sec_first=%d sec_count=%d src_row0=%lld)", d->sec_first,
                    d->sec_count, (long long)d->src_row0);

    // ---- section loop (ref preproc.h:379-408) -> fast warp-tiles + generic tiles; cached on the geometry and the map
    const bool fast_ok = ctx->mss_fast != 0 && (((uintptr_t)d_mss & 15) == 0) && ((d->pitch_px * 2) % 16 == 0) && d->wb >= 16;
    struct Key { int wb, lps, overlap, keep, min_lines, fast, tile_rows, sec_first, sec_count; int64_t lines, off, src_row0; double cX[8], cY[12]; } key{};
    key.wb = d->wb; key.lps = lps; key.overlap = overlap; key.keep = d->keep_leading != 0; key.min_lines = min_lines;
    // long strips: taller warp-tiles amortise the per-tile prologue (C3: 501 -> 515 Gpixel/s, tools/bench_mss.py)
    const int tile_rows = (ctx->mss_fast_rows == 128 && d->lines >= 32768) ? 256 : ctx->mss_fast_rows;
    key.fast = fast_ok; key.tile_rows = tile_rows; key.lines = d->lines; key.off = d->line_offset;
    key.sec_first = d->sec_first; key.sec_count = d->sec_count; key.src_row0 = d->src_row0;
    memcpy(key.cX, d->cX, sizeof key.cX); memcpy(key.cY, d->cY, sizeof key.cY);
    const uint8_t *kbytes = reinterpret_cast<const uint8_t *>(&key);
    if (ctx->mss_plan_key.size() != sizeof key || memcmp(ctx->mss_plan_key.data(), kbytes, sizeof key) != 0) {
        std::vector<mssfast::Section> secs;
        const int64_t processed = build_sections(d, secs);
        for (const mssfast::Section &sc : secs)
            if (sc.sec_off < 0)
                return fail(OIP_E_RANGE, "oip_band_align_merge: section shard starts at strip line %lld, before src_row0 = %lld",
                            (long long)(sc.sec_off + d->src_row0), (long long)d->src_row0);
        std::vector<mss::Tile> tiles;
        std::vector<mssfast::FTile> ftiles;
        mssfast::plan(d, secs, fast_ok, tile_rows, tiles, ftiles);
        ctx->mss_plan_rows = processed;
        float tab[132] = {};
        oip_cubic_tab(tab);
        tab[128] = tab[129] = -0.0f; // run-time (-0.0,-0.0) addend of the packed products (oip_common.cuh mul2)
        const size_t fast_off = (1024 + tiles.size() * sizeof(mss::Tile) + 63) / 64 * 64;
        const size_t bytes = fast_off + ftiles.size() * sizeof(mssfast::FTile) + 64;
        if (bytes > ctx->d_mss_plan_cap) {
            if (ctx->d_mss_plan) { OIP_CUDA(cudaStreamSynchronize(ctx->stream)); OIP_CUDA(cudaFree(ctx->d_mss_plan)); ctx->d_mss_plan = nullptr; }
            OIP_CUDA(cudaMalloc(&ctx->d_mss_plan, bytes * 2));
            ctx->d_mss_plan_cap = bytes * 2;
        }
        OIP_CUDA(cudaMemcpyAsync(ctx->d_mss_plan, tab, sizeof tab, cudaMemcpyHostToDevice, ctx->stream));
        if (!tiles.empty())
            OIP_CUDA(cudaMemcpyAsync((uint8_t *)ctx->d_mss_plan + 1024, tiles.data(), tiles.size() * sizeof(mss::Tile), cudaMemcpyHostToDevice, ctx->stream));
        if (!ftiles.empty())
            OIP_CUDA(cudaMemcpyAsync((uint8_t *)ctx->d_mss_plan + fast_off, ftiles.data(), ftiles.size() * sizeof(mssfast::FTile), cudaMemcpyHostToDevice, ctx->stream));
        OIP_CUDA(cudaStreamSynchronize(ctx->stream)); // the host vectors go out of scope
        ctx->mss_plan_key.assign(kbytes, kbytes + sizeof key);
        ctx->mss_plan_tiles = (int64_t)tiles.size();
        ctx->mss_fast_ctas = (int64_t)(ftiles.size() / mssfast::WARPS);
        ctx->mss_fast_off = fast_off;
    }
    if (rows_written) *rows_written = ctx->mss_plan_rows;
    if (ctx->mss_plan_tiles == 0 && ctx->mss_fast_ctas == 0) return OIP_OK;

    // ---- fast kernel on the caller's stream; the generic tiles run next to it on a side stream
    const bool fork = ctx->mss_fast_ctas > 0 && ctx->mss_plan_tiles > 0;
    if (fork) {
        if (!ctx->aux_stream) {
            OIP_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            OIP_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            OIP_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        OIP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
        OIP_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    }
    cudaStream_t gstream = fork ? ctx->aux_stream : ctx->stream;

    mss::Params P{};
    P.base = (const uint8_t *)d_mss;
    P.pitch_bytes = d->pitch_px * 2;
    for (int b = 0; b < 4; ++b) P.kb[b] = d->d_kb[b];
    for (int i = 0; i < 8; ++i) P.cX[i] = d->cX[i];
    for (int i = 0; i < 12; ++i) P.cY[i] = d->cY[i];
    P.tab = (const float *)ctx->d_mss_plan;
    P.tiles = (const mss::Tile *)((const uint8_t *)ctx->d_mss_plan + 1024);
    P.out = d_out; P.err = ctx->d_err; P.fmt = d->fmt; P.wb = d->wb;
    P.vec_ok = (((uintptr_t)d_mss & 3) == 0 && (P.pitch_bytes & 3) == 0 && (d->wb % 2) == 0) ? 1 : 0;
    const size_t smem = (size_t)mss::RING * mss::SWC * 4;
    if (!ctx->mss_attr_set) {
        OIP_CUDA(cudaFuncSetAttribute(mss::band_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->mss_attr_set = true;
    }
    if (ctx->mss_plan_tiles > 0) {
        mss::band_align_kernel<<<(unsigned)ctx->mss_plan_tiles, mss::NT, smem, gstream>>>(P);
        OIP_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    if (ctx->mss_fast_ctas > 0) {
        mssfast::Params F;
        memset(&F, 0, sizeof F);
        int rc = mssfast::encode(&F.tmap, d_mss, 4 * d->wb, d->lines - d->src_row0, d->pitch_px * 2);
        if (rc) return rc;
        for (int b = 0; b < 4; ++b) F.kb[b] = d->d_kb[b];
        memcpy(F.cX, d->cX, sizeof F.cX); memcpy(F.cY, d->cY, sizeof F.cY);
        F.tiles = (const mssfast::FTile *)((const uint8_t *)ctx->d_mss_plan + ctx->mss_fast_off);
        F.out = d_out; F.tab = (const float *)ctx->d_mss_plan; F.nz = 0x8000000080000000ull; F.wb = d->wb; F.swap = d->fmt == OIP_FMT_BE16;
        F.n_stage = std::max(2, std::min(8, ctx->pan_fast_stages));
        rc = mssfast::launch(ctx, F, ctx->mss_fast_ctas);
        if (rc) return rc;
    }
    if (fork) {
        OIP_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
        OIP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    // (no device-side error channel any more: a polynomial that exceeds the staged window is resampled tap by tap inside
    // band_align_kernel, so the call returns without a device-to-host copy or a stream synchronisation)
    return OIP_OK;
}
