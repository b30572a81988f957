// mss_fast.cu -- regular-interior fast path of the MSS band alignment (band split + per-band RRC + per-band polynomial
// bicubic remap + 4-channel merge).  Replaces the same reference code as oip_band_align_merge
// (ref preproc.h:56-80 split, :202-222 RRC, :351-468 DoInterBandAlignment + cv::remap + cv::merge).
//
// The reference evaluates a polynomial map per pixel (x linear, y quadratic in x, ref preproc.h:443-449).  Its
// fixed-point form (OpenCV: cvRound(map*32)) has, for one band and one column, a constant sub-pixel phase
// (fx, fy) and a constant row offset over long runs of rows; over runs of columns the source column offset and the
// row offset are constant too.  The host planner below PROVES this per column with the reference's own arithmetic
// and cuts such regions into warp-tiles; everything else (rows where y + c(x) crosses a power of two and the float
// rounding of the map changes, section borders, columns near the band's edge) stays on the exact generic kernel of
// mss.cu.
//
// Warp-tile: one band, 2 x NH output columns (a lane owns one column of the left and one of the right half, packed
// as an FP32 pair), n_rows output rows.  Same machinery as pan_fast.cu: 2-D TMA tensor stages per warp, conversion on
// the FP64 / conversion pipes, the 3 extra window columns by SHFL from lanes +1..+3, scatter-form bicubic sum with
// rotating accumulators, OpenCV's interior accumulation order, no FMA contraction.  The weights are per lane
// (each column has its own phase), held as 16 packed pairs.
#include <algorithm>
#include <cmath>

#include "mss_plan.cuh"

namespace oip {
namespace mssfast {
using namespace tmaw;

constexpr int NH = 29;           // output columns per half (lanes 29..31 only supply window columns)
constexpr int RS = 8;            // source rows per TMA stage
constexpr int BOX_W32 = 36;      // 72 samples per box row: 2*NH + 3 window columns + up to 7 of origin alignment
constexpr int ROW_BYTES = BOX_W32 * 4;      // 144
constexpr int STAGE_BYTES = ROW_BYTES * RS; // 1152 = 9 x 128
constexpr int MAX_STAGE = 8;
static_assert(STAGE_BYTES % 128 == 0, "stage alignment");

// ---- the reference's map arithmetic (ref preproc.h:443-449): fp64, left to right, then float, then OpenCV's 1/32 grid
__device__ __forceinline__ int dev_sx(const double *cX, int b, int x)
{
    const double dxx = (double)(x * 4);
    const float m = __double2float_rn(__ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(cX[2 * b + 1], dxx), cX[2 * b]), dxx), 4.0));
    return __float2int_rn(__fmul_rn(m, 32.0f));
}
__device__ __forceinline__ int dev_sy(const double *cY, int b, int x, int y)
{
    const double dxx = (double)(x * 4);
    const double Ay = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(cY[3 * b + 2], dxx), dxx), __dmul_rn(cY[3 * b + 1], dxx)), cY[3 * b]);
    const float m = __double2float_rn(__ddiv_rn(__dadd_rn(Ay, (double)((int64_t)y * 4)), 4.0));
    return __float2int_rn(__fmul_rn(m, 32.0f));
}
static int host_sx(const double *cX, int b, int x)
{
    const double dxx = (double)(x * 4);
    const volatile double t0 = cX[2 * b + 1] * dxx;
    const volatile double t1 = t0 + cX[2 * b];
    const volatile double t2 = t1 + dxx;
    const float m = (float)(t2 / 4.0);
    const volatile float v = m * 32.0f;
    return (int)lrintf(v);
}
static double host_Ay(const double *cY, int b, int x)
{
    const double dxx = (double)(x * 4);
    const volatile double q0 = cY[3 * b + 2] * dxx;
    const volatile double q1 = q0 * dxx;
    const volatile double l1 = cY[3 * b + 1] * dxx;
    const volatile double s0 = q1 + l1;
    const volatile double s1 = s0 + cY[3 * b];
    return s1;
}
static int host_sy(double Ay, int y)
{
    const volatile double t = Ay + (double)((int64_t)y * 4);
    const float m = (float)(t / 4.0);
    const volatile float v = m * 32.0f;
    return (int)lrintf(v);
}
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

struct WarpCtx {
    const CUtensorMap *tm;
    uint32_t stage0, bar0;
    int ns, lane;
};
__device__ __forceinline__ void issue_stage(const WarpCtx &C, int slot, int x, int64_t y)
{
    const uint32_t bar = C.bar0 + 8u * slot, dst = C.stage0 + (uint32_t)slot * STAGE_BYTES;
    mbar_expect_tx_u32(bar, STAGE_BYTES);
    tma_load_2d(dst, C.tm, x >> 1, (int)y, bar);
}
__device__ __forceinline__ f2 shfl_down_f2(f2 v, int d)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_down_sync(0xffffffffu, lo, d);
    hi = __shfl_down_sync(0xffffffffu, hi, d);
    return ((f2)hi << 32) | lo;
}

// EDGE: the tile touches the band's left / right border: source columns outside [0, wb) are zero (cv::remap's
// BORDER_CONSTANT on the band plane; in the mixed line they hold the neighbour band) and output columns whose 4x4
// footprint leaves the band use OpenCV's border accumulation order (one flat chain from 0).  Both sums
// are computed from the same products and selected per column.
template <int MODE, bool SWAP, bool EDGE>
__device__ __forceinline__ void mss_tile(const Params &P, const FTile &T, const WarpCtx &C)
{
    const int lane = C.lane, ns = C.ns, b = T.band, wb = P.wb;
    const int n_chunks = (T.n_rows + 3 + RS - 1) / RS;
    const int line_col = b * wb + T.ix0;       // column of the first window sample in the mixed MSS line
    const int x0 = line_col & ~7;              // box origin: 16-byte boundary of the tensor row
    const uint32_t offL = 2u * (uint32_t)(line_col - x0 + lane), offR = offL + 2u * (uint32_t)T.nh;
    if (lane == 0) {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage(C, c, x0, T.src_row0 + c * RS);
    }
    // (k,b) of the two detectors this lane converts
    const int rawL = T.ix0 + lane, rawR = T.ix0 + T.nh + lane;
    const int colL = max(0, min(rawL, wb - 1)), colR = max(0, min(rawR, wb - 1));
    const bool inL = !EDGE || (rawL >= 0 && rawL < wb), inR = !EDGE || (rawR >= 0 && rawR < wb);
    const bool bordL = EDGE && !(rawL >= 0 && rawL <= wb - 4), bordR = EDGE && !(rawR >= 0 && rawR <= wb - 4);
    double kL = 1.0, bL = 0.0, kR = 1.0, bR = 0.0;
    if (MODE != 0) {
        const double *kbp = P.kb[b];
        kL = kbp[2 * colL]; bL = kbp[2 * colL + 1];
        kR = kbp[2 * colR]; bR = kbp[2 * colR + 1];
    }
    // per-lane weights: w[r][c] = fl32(wy[r] * wx[c]) of my left / right output column
    const int xl = max(0, min(T.x_begin + lane, wb - 1)), xr = max(0, min(T.x_begin + T.nh + lane, wb - 1));
    const int fxl = dev_sx(P.cX, b, xl) & 31, fxr = dev_sx(P.cX, b, xr) & 31;
    const int fyl = dev_sy(P.cY, b, xl, T.ya) & 31, fyr = dev_sy(P.cY, b, xr, T.ya) & 31;
    const f2 nz = P.nz; // (measured: 0.396 ms against 0.407 ms with the pair loaded from the plan header)
    f2 W[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            W[r][c] = pk(__fmul_rn(__ldg(P.tab + 4 * fyl + r), __ldg(P.tab + 4 * fxl + c)),
                         __fmul_rn(__ldg(P.tab + 4 * fyr + r), __ldg(P.tab + 4 * fxr + c)));
    const bool actL = lane < T.nh, actR = lane < T.n_right;
    const int64_t pitch = (int64_t)wb * 4;
    uint16_t *oL = P.out + T.out_off + 4 * lane - 3 * pitch; // output row (m - 3) while source row m is consumed
    uint16_t *oR = oL + 4 * T.nh;
    const int n_rows = T.n_rows;
    f2 A1 = 0ull, A2 = 0ull, A3 = 0ull, F1 = 0ull, F2 = 0ull, F3 = 0ull;

    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        const uint32_t sa = C.stage0 + (uint32_t)slot * STAGE_BYTES;
#pragma unroll
        for (int rr = 0; rr < RS; ++rr) {
            uint32_t sL, sR;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sL) : "r"(sa + rr * ROW_BYTES + offL));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sR) : "r"(sa + rr * ROW_BYTES + offR));
            if (SWAP) { sL = __byte_perm(sL, 0u, 0x4401); sR = __byte_perm(sR, 0u, 0x4401); }
            float fl, fr;
            if (MODE == 0) {
                fl = (float)(uint16_t)sL;
                fr = (float)(uint16_t)sR;
            } else {
                fl = (float)(uint16_t)rrc_d<MODE>(__uint2double_rn(sL), kL, bL);
                fr = (float)(uint16_t)rrc_d<MODE>(__uint2double_rn(sR), kR, bR);
            }
            if (EDGE) {
                fl = inL ? fl : 0.f;
                fr = inR ? fr : 0.f;
            }
            f2 win[4];
            win[0] = pk(fl, fr);
#pragma unroll
            for (int j = 1; j < 4; ++j) win[j] = shfl_down_f2(win[0], j);
            // per row ((s0*w0 + s1*w1) + s2*w2) + s3*w3, rows accumulated in order 0,1,2,3 (OpenCV interior order)
            f2 d[4], f[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const f2 p0 = mul2(win[0], W[r][0], nz), p1 = mul2(win[1], W[r][1], nz), p2 = mul2(win[2], W[r][2], nz), p3 = mul2(win[3], W[r][3], nz);
                d[r] = add2(add2(add2(p0, p1), p2), p3);
                if (EDGE) {
                    const f2 prev = r == 0 ? 0ull : (r == 1 ? F1 : (r == 2 ? F2 : F3));
                    f[r] = add2(add2(add2(add2(prev, p0), p1), p2), p3);
                }
            }
            f2 out = add2(A3, d[3]);
            A3 = add2(A2, d[2]);
            A2 = add2(A1, d[1]);
            A1 = d[0];
            if (EDGE) {
                out = pk(bordL ? lo_of(f[3]) : lo_of(out), bordR ? hi_of(f[3]) : hi_of(out));
                F3 = f[2]; F2 = f[1]; F1 = f[0];
            }
            const int m = c * RS + rr;
            const bool rows_ok = (unsigned)(m - 3) < (unsigned)n_rows;
            if (rows_ok && actL) *oL = (uint16_t)cast_u16(lo_of(out));
            if (rows_ok && actR) *oR = (uint16_t)cast_u16(hi_of(out));
            oL += pitch;
            oR += pitch;
        }
        __syncwarp();
        if (lane == 0 && c + ns < n_chunks) issue_stage(C, slot, x0, T.src_row0 + (int64_t)(c + ns) * RS);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
    }
}

__global__ void __launch_bounds__(WARPS * 32, 4) mss_fast_kernel(const __grid_constant__ Params P)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[WARPS][MAX_STAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FTile T = P.tiles[(int64_t)blockIdx.x * WARPS + warp];
    if (T.band < 0) return;
    WarpCtx C;
    C.ns = P.n_stage;
    C.lane = lane;
    C.tm = &P.tmap;
    C.stage0 = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(warp * C.ns) * STAGE_BYTES;
    C.bar0 = smem_u32(&bars[warp][0]);
    if (lane == 0) {
        for (int s = 0; s < C.ns; ++s) mbar_init_u32(C.bar0 + 8u * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    // RRC mode of the warp (see pan_fast.cu): decided from the coefficients, never from the data
    int mode = 0;
    if (P.kb[T.band]) {
        const double *kbp = P.kb[T.band];
        const int cl = max(0, min(T.ix0 + lane, P.wb - 1)), cr = max(0, min(T.ix0 + T.nh + lane, P.wb - 1));
        bool general = false;
        const int cols[2] = {cl, cr};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double k = kbp[2 * cols[j]], bb = kbp[2 * cols[j] + 1];
            general = general || !(k >= 0.0 && bb >= 0.0 && __dadd_rn(__dmul_rn(k, 65535.0), bb) < 2147483648.0);
        }
        mode = __any_sync(0xffffffffu, general) ? 2 : 1;
    }
    const bool swap = P.swap != 0;
#define OIP_MSS_GO(M)                                                   \
    do {                                                                \
        if (T.pad) {                                                    \
            if (swap) mss_tile<M, true, true>(P, T, C);                 \
            else mss_tile<M, false, true>(P, T, C);                     \
        } else {                                                        \
            if (swap) mss_tile<M, true, false>(P, T, C);                \
            else mss_tile<M, false, false>(P, T, C);                    \
        }                                                               \
    } while (0)
    if (mode == 1) OIP_MSS_GO(1);
    else if (mode == 0) OIP_MSS_GO(0);
    else OIP_MSS_GO(2);
#undef OIP_MSS_GO
}

// ------------------------------------------------------------------------------------------------ host planner
static void emit_generic(std::vector<mss::Tile> &tiles, const Section &S, int band, int xa, int xb, int ya, int yb)
{
    const int TW = 240, TH = 256; // mss.cu tile limits
    for (int y = ya; y < yb; y += TH)
        for (int x = xa; x < xb; x += TW) {
            mss::Tile t{};
            t.band = band; t.x_begin = x; t.x_end = std::min(xb, x + TW);
            t.rows = S.rows; t.y0 = y; t.n_rows = std::min(TH, yb - y);
            t.sec_off = S.sec_off; t.dst_row0 = S.dst_row0 + (y - S.y0);
            tiles.push_back(t);
        }
}

// Splits every section x band into fast warp-tiles and generic tiles.  fast == false: everything generic.
void plan(const oip_mss_desc *d, const std::vector<Section> &secs, bool fast, int tile_rows, std::vector<mss::Tile> &tiles,
          std::vector<FTile> &ftiles)
{
    const int wb = d->wb;
    std::vector<int> sx(wb), ixo(wb);
    std::vector<double> Ay(wb);
    std::vector<int> sya(wb), D(wb);
    std::vector<char> ok(wb);
    std::vector<FTile> fl;
    for (const Section &S : secs) {
        for (int b = 0; b < 4; ++b) {
            if (!fast || wb < 16) { emit_generic(tiles, S, b, 0, wb, S.y0, S.rows); continue; }
            double cmin = 1e300, cmax = -1e300;
            for (int x = 0; x < wb; ++x) {
                sx[x] = host_sx(d->cX, b, x);
                ixo[x] = sat_short(sx[x] >> 5) - 1 - x;
                Ay[x] = host_Ay(d->cY, b, x);
                cmin = std::min(cmin, Ay[x] / 4.0);
                cmax = std::max(cmax, Ay[x] / 4.0);
            }
            if (!(cmin > -30000.0 && cmax < 30000.0)) { emit_generic(tiles, S, b, 0, wb, S.y0, S.rows); continue; }
            // rows where y + c(x) may cross a power of two for some column: the float rounding of the map changes there
            std::vector<std::pair<int, int>> zones; // [a, b) section-local rows, sorted
            for (int k = 0; k <= 16; ++k) {
                const int p = k == 0 ? 0 : (1 << k);
                zones.push_back({p - (int)std::ceil(cmax) - 2, p - (int)std::floor(cmin) + 3});
            }
            // rows whose footprint can leave the section Mat at its top / bottom: border order, generic
            zones.push_back({-(1 << 30), 3 - (int)std::floor(cmin)});
            zones.push_back({S.rows - 6 - (int)std::ceil(cmax), 1 << 30});
            std::sort(zones.begin(), zones.end());
            int y = S.y0;
            size_t zi = 0;
            while (y < S.rows) {
                while (zi < zones.size() && zones[zi].second <= y) ++zi;
                if (zi < zones.size() && zones[zi].first <= y) { // inside a zone: generic up to its end
                    const int ye = std::min(S.rows, zones[zi].second);
                    emit_generic(tiles, S, b, 0, wb, y, ye);
                    y = ye;
                    continue;
                }
                const int yb_run = std::min(S.rows, zi < zones.size() ? zones[zi].first : S.rows);
                // run [y, yb_run): tile rows, verify every column at both ends of each tile
                const int len = yb_run - y;
                const int n_t = (len + tile_rows - 1) / tile_rows;
                const int h = (len + n_t - 1) / n_t;
                for (int ya = y; ya < yb_run; ya += h) {
                    const int ye = std::min(yb_run, ya + h);
                    for (int x = 0; x < wb; ++x) {
                        sya[x] = host_sy(Ay[x], ya);
                        const int sye = host_sy(Ay[x], ye - 1);
                        D[x] = sat_short(sya[x] >> 5) - 1 - ya;
                        const int ix = ixo[x] + x;
                        ok[x] = sye == sya[x] + 32 * (ye - 1 - ya) && ix > -30000 && ix < 30000 && ya + D[x] >= 0 &&
                                (ye - 1) + D[x] + 3 <= S.rows - 1;
                    }
                    int x = 0;
                    while (x < wb) {
                        if (!ok[x]) {
                            int xe = x + 1;
                            while (xe < wb && !ok[xe]) ++xe;
                            emit_generic(tiles, S, b, x, xe, ya, ye);
                            x = xe;
                            continue;
                        }
                        int xe = x + 1;
                        while (xe < wb && ok[xe] && ixo[xe] == ixo[x] && D[xe] == D[x]) ++xe;
                        // strips of <= 2*NH columns, equal widths
                        const int width = xe - x, n_s = (width + 2 * NH - 1) / (2 * NH);
                        int xs = x;
                        for (int s = 0; s < n_s; ++s) {
                            const int wS = width / n_s + (s < width % n_s ? 1 : 0);
                            FTile t{};
                            t.band = b; t.x_begin = xs; t.nh = (wS + 1) / 2; t.n_right = wS - t.nh;
                            t.ix0 = ixo[x] + xs; t.ya = ya; t.n_rows = ye - ya;
                            // pad = 1: EDGE variant (some footprint of the strip, window lanes included, leaves the band)
                            t.pad = (t.ix0 < 0 || t.ix0 + 2 * t.nh + 2 > wb - 1) ? 1 : 0;
                            t.src_row0 = S.sec_off + ya + D[x];
                            t.out_off = ((S.dst_row0 + (ya - S.y0)) * wb + xs) * 4 + b;
                            fl.push_back(t);
                            xs += wS;
                        }
                        x = xe;
                    }
                }
                y = yb_run;
            }
        }
    }
    // raster order: row band, then columns, the 4 bands of a block adjacent (they fill the same 8-byte pixels)
    const int64_t row_elems = (int64_t)wb * 4;
    std::stable_sort(fl.begin(), fl.end(), [&](const FTile &a, const FTile &c) {
        const int64_t ra = a.out_off / row_elems / tile_rows, rc = c.out_off / row_elems / tile_rows;
        if (ra != rc) return ra < rc;
        const int ca = a.x_begin / (2 * NH), cc = c.x_begin / (2 * NH);
        if (ca != cc) return ca < cc;
        return a.band < c.band;
    });
    ftiles = fl;
    FTile none{};
    none.band = -1;
    while (ftiles.size() % WARPS) ftiles.push_back(none);
}

int launch(oip_ctx *ctx, const Params &P, int64_t n_ctas)
{
    static bool attr_set = false;
    if (!attr_set) {
        OIP_CUDA(cudaFuncSetAttribute(mss_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * MAX_STAGE * STAGE_BYTES + 128));
        attr_set = true;
    }
    const size_t smem = (size_t)WARPS * P.n_stage * STAGE_BYTES + 128;
    mss_fast_kernel<<<(unsigned)n_ctas, WARPS * 32, smem, ctx->stream>>>(P);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

int encode(CUtensorMap *tm, const void *base, int line_px, int64_t lines, int64_t pitch_bytes)
{
    return tmaw::encode_tmap_u32(tm, base, line_px, lines, pitch_bytes, BOX_W32, RS);
}

} // namespace mssfast
} // namespace oip
