// pan_fast.cu -- the regular-interior fast path of the fused PAN pipeline (unpack -> RRC -> cubic shift -> concat).
//
// Everything the planner can prove regular (footprints inside the section's fresh rows and the
// CCD's columns, fixed-point map advancing by exactly one source pixel per output pixel) runs here; section
// edges, image borders, map-rounding anomalies, packed / tiled / unaligned inputs stay on the exact generic
// kernel of pan_pipeline.cu.  Replaces the same reference code as oip_pan_pipeline (ref imageop.h:129-138,
// :230-275, :277-363, stitcher.h:83-139).
//
// Structure: a CTA is four INDEPENDENT warps, each marching down its own warp-tile; there is no shared ring
// and no __syncthreads.
//   load      one lane issues 2-D TMA tensor copies (cp.async.bulk.tensor, SASS UTMALDG) of RC raw rows into
//             the warp's private stage ring, n_stage deep, each stage behind its own mbarrier
//   convert   a lane owns 4 + 4 source columns (left half / right half of the tile): byte swap by PRMT, fp64
//             RRC with (k,b) in registers (magic-number int<->double, no conversion-pipe op), one I2F.U16
//   exchange  the 3 extra window columns of each half come from lane+1 by SHFL (no shared-memory ring)
//   resample  scatter form of OpenCV's bicubic sum: a source row is multiplied once into the 4 output rows
//             it feeds (accumulators rotate through registers), products and sums in OpenCV's own order,
//             packed FFMA2/FADD2 over the (left, right) pixel pair, no FMA contraction
//   store     F2I.U16 (round-half-even, saturating) + 64-bit stores straight into the trimmed output raster
// COPY warp-tiles (unshifted CCDs) use the same staging: swap + RRC + 128-bit stores.
#include "pan_fast.cuh"
#include "tma_warp.cuh"

#ifndef OIP_DBG_VARIANT
#define OIP_DBG_VARIANT 0 // 1..3: timing experiments that drop parts of the row loop (wrong output; tools/probes/build_variant.py)
#endif

namespace oip {
namespace panfast {

constexpr int ROW_BYTES = BOX_W * 4;                          // 544: a box row is BOX_W 32-bit elements = 272 samples
constexpr int STAGE_BYTES = ROW_BYTES * RC;                   // 2176 = 17 x 128: TMA destinations are 128-byte aligned
constexpr int MAX_STAGE = 8;
static_assert(RC == 4, "the row loop is unrolled by the 4-deep accumulator rotation");
static_assert(STAGE_BYTES % 128 == 0, "stage alignment");

using namespace tmaw;

// 4 consecutive samples that start DM halfwords into the aligned 8-byte shared-memory word at `a`, as floats
// (after byte swap and RRC).  The TMA unit only accepts box origins on 16-byte boundaries of a tensor row
// (measured, tools/probes/tma_probe.cu: any other coordinate raises "illegal instruction"), so the source window starts
// (src_x0 & 7) samples into the box; DM = that offset mod 4 is a template parameter.
// The kernel is bound by instruction issue (a packed FFMA2/FADD2 holds the issue port of its SM sub-partition for two
// cycles, every other instruction for one; tools/probes/mix_rates.cu), so every step here is the form with the fewest
// instructions: I2F.U16 takes the low 16 bits of the RRC result (the reference's mod-2^16 wrap) or a halfword of
// the raw word directly; the conversion pipe (one warp instruction per 8 cycles) has the headroom.
template <int MODE, int DM, bool SWAP>
__device__ __forceinline__ void convert4(uint32_t a, const double *k, const double *b, float *f)
{
    constexpr int NW = (DM & 1) ? 3 : 2;
    uint32_t w[3] = {0u, 0u, 0u};
    if (DM == 0) { const uint2 A = lds64(a); w[0] = A.x; w[1] = A.y; }
    else if (DM == 1) { const uint2 A = lds64(a); w[0] = A.x; w[1] = A.y; w[2] = lds32(a + 8); }
    else if (DM == 2) { w[0] = lds32(a + 4); w[1] = lds32(a + 8); }
    else { w[0] = lds32(a + 4); const uint2 B = lds64(a + 8); w[1] = B.x; w[2] = B.y; }
    if (SWAP) {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = __byte_perm(w[i], 0u, 0x2301);
    }
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { // I2F.U16 with a halfword selector
            const int h = j + (DM & 1);
            f[j] = (h & 1) ? (float)(uint16_t)(w[h >> 1] >> 16) : (float)(uint16_t)(w[h >> 1] & 0xFFFFu);
        }
    } else {
        D2 d[3];
#pragma unroll
        for (int i = 0; i < NW; ++i) d[i] = split_word(w[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int h = j + (DM & 1);
            const double sd = (h & 1) ? d[h >> 1].hi : d[h >> 1].lo;
            f[j] = (float)(uint16_t)rrc_d<MODE>(sd, k[j], b[j]);
        }
    }
}

struct WarpCtx {
    const CUtensorMap *tm;
    uint32_t stage0, bar0; // shared-memory addresses of this warp's stage ring and barriers
    uint32_t le0;          // packed sources: the warp's unpacked (u16 LE) image of the stage being consumed
    int ns, lane;
};

__device__ __forceinline__ uint32_t fdiv(const FastDiv &d, uint32_t n)
{
    const uint32_t t = __umulhi(d.m, n);
    return (t + ((n - t) >> d.s1)) >> d.s2;
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// Stage loader for OIP_FMT_BE16_TILES sources: the same RC x BOX_W-word stage image the tensor copy produces, gathered
// straight from the sub-images of the IMDT stream, so the frame tiles never become a raster in HBM (ref
// aux_separator.h:341-372 MergeSubImage writes that raster).  Frames sit at arbitrary 4-byte offsets of
// the stream (a frame is 12 mod 16 bytes long), which rules the TMA unit out (16-byte global alignment): every lane
// copies 17 words of ONE row with 4-byte cp.async (LDGSTS, zero fill by src-size 0) and the stage's mbarrier counts the
// 32 lanes' completions (cp.async.mbarrier.arrive.noinc).  No loader state lives across the row loop: the row's frame /
// tile row / line come from its index by two multiply-high divisions.
// x: first sample of the box (multiple of 8, may be < 0 for EDGE tiles), y: first row, c0 = floor(x / tile_cols)
__device__ __forceinline__ void issue_stage_tiled(const FastCcd &S, const WarpCtx &C, int slot, int x, int y, int c0)
{
    const int rr = C.lane & 3, kq = C.lane >> 2;
    const uint32_t bar = C.bar0 + 8u * slot;
    const uint32_t dst = C.stage0 + (uint32_t)slot * STAGE_BYTES + (uint32_t)rr * ROW_BYTES + 4u * (uint32_t)kq;
    const uint32_t yr = (uint32_t)(y + rr);
    const uint32_t f = fdiv(S.div_lpf, yr), rl = yr - f * S.div_lpf.d;
    const uint32_t r = fdiv(S.div_tl, rl), line = rl - r * S.div_tl.d;
    const bool row_ok = f < (uint32_t)S.n_frames;          // rows past the last frame (chunk rounding): zero fill
    const int xa = c0 * S.tile_cols, xb = xa + S.tile_cols;
    const int xl = x + 2 * kq;                             // sample column of my first word
    const int64_t *tab = S.tile_off + ((int64_t)f * 40 + r * 8);
    int64_t oA = -1, oB = -1;
    if (row_ok && (unsigned)c0 < 8u) oA = __ldg(tab + c0);
    const uint8_t *pA = S.tile_base;
    uint32_t szA = 0;
    if (oA >= 0) { pA += oA + 2 * ((int64_t)line * S.tile_cols + (xl - xa)); szA = 4; }
    if (x >= 0 && x + 2 * BOX_W <= xb) {                   // (warp-uniform) the whole window lies in one sub-image column
#pragma unroll
        for (int j = 0; j < BOX_W / 8; ++j) cp_async4(dst + 32u * j, pA + 32 * j, szA);
    } else {
        if (row_ok && (unsigned)(c0 + 1) < 8u) oB = __ldg(tab + c0 + 1);
        const uint8_t *pB = S.tile_base;
        uint32_t szB = 0;
        if (oB >= 0) { pB += oB + 2 * ((int64_t)line * S.tile_cols + (xl - xb)); szB = 4; }
        const int jb = xl >= xb ? 0 : (xb - xl + 15) >> 4; // first word of mine that lies in column c0 + 1
#pragma unroll
        for (int j = 0; j < BOX_W / 8; ++j) {
            const bool b = j >= jb;
            cp_async4(dst + 32u * j, (b ? pB : pA) + 32 * j, b ? szB : szA);
        }
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// warp-collective.  Line formats: one lane issues the 2-D tensor copy (x: first sample of the box, a multiple of 8; the
// tensor map counts 32-bit elements).  Tiled sources: see above.
// Packed sources (PB = 12 / 10, MSB-first bit stream per line): the box origin must sit on a 16-byte boundary of the
// packed line, i.e. on a multiple of 32 (12-bit: 48 bytes) / 64 (10-bit: 80 bytes) samples; the box covers the 16-sample
// groups that hold the 272-sample window.
template <int PB> __device__ __forceinline__ int packed_origin(int x) { return PB == 12 ? (x & ~31) : (x & ~63); }
template <bool TILED, int PB = 0>
__device__ __forceinline__ void issue_stage(const FastParams &P, const FastTile &T, const WarpCtx &C, int slot, int x, int y)
{
    if (TILED) {
        issue_stage_tiled(P.ccd[T.ccd], C, slot, x, y, T.tmap);
    } else if (C.lane == 0) {
        const uint32_t bar = C.bar0 + 8u * slot, dst = C.stage0 + (uint32_t)slot * STAGE_BYTES;
        if (PB == 0) {
            mbar_expect_tx_u32(bar, STAGE_BYTES);
            tma_load_2d(dst, C.tm, x >> 1, y, bar);
        } else {
            constexpr int BOXW = PB == 12 ? PBOX12 : PBOX10;
            mbar_expect_tx_u32(bar, (uint32_t)(BOXW * 4 * RC));
            tma_load_2d(dst, C.tm, (packed_origin<PB>(x) * PB) >> 5, y, bar);
        }
    }
}
// Packed sources: the stage that just arrived (RC rows of packed bytes) -> the warp's u16 little-endian stage image, the
// same layout a tensor copy of 16-bit lines produces (ROW_BYTES per row, sample x at 2 * (x - x_le)).  16 samples = 6 / 5
// aligned words rebuilt big-endian (one PRMT each), one funnel shift + one shift per sample, two 16-byte stores; 18 groups
// per row cover the window, 72 per stage = 2.25 per lane.  (oip_unpack_lines does the same between two HBM buffers.)
template <int PB>
__device__ __forceinline__ void unpack_stage(const WarpCtx &C, uint32_t stage, int x_le)
{
    constexpr int NW = PB * 16 / 32, BOXB = (PB == 12 ? PBOX12 : PBOX10) * 4;
    const int xg = x_le & ~15, xp = packed_origin<PB>(x_le);
    const uint32_t g_byte0 = (uint32_t)(((xg - xp) * PB) >> 3);     // first group's offset inside a packed box row
    const int lead = x_le - xg;                                      // 0 or 8: samples of group 0 in front of the image
#pragma unroll 1
    for (int it = C.lane; it < 18 * RC; it += 32) {
        const int rr = it / 18, g = it - rr * 18;
        const uint32_t src = stage + (uint32_t)rr * BOXB + g_byte0 + (uint32_t)(g * NW * 4);
        uint32_t be[NW + 1];
#pragma unroll
        for (int k = 0; k < NW; ++k) be[k] = __byte_perm(lds32(src + 4u * k), 0u, 0x0123);
        be[NW] = 0u;
        uint32_t px[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int bit = PB * k;
            px[k] = __funnelshift_l(be[(bit >> 5) + 1], be[bit >> 5], bit & 31) >> (32 - PB);
        }
        const int s0 = 16 * g - lead;                                // image sample index of the group's first sample
        const uint32_t dst = C.le0 + (uint32_t)rr * ROW_BYTES + (uint32_t)(2 * s0);
        if (s0 >= 0 && s0 + 8 <= 2 * BOX_W)
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(px[0] | (px[1] << 16)), "r"(px[2] | (px[3] << 16)),
                         "r"(px[4] | (px[5] << 16)), "r"(px[6] | (px[7] << 16)) : "memory");
        if (s0 + 8 >= 0 && s0 + 16 <= 2 * BOX_W)
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16u), "r"(px[8] | (px[9] << 16)), "r"(px[10] | (px[11] << 16)),
                         "r"(px[12] | (px[13] << 16)), "r"(px[14] | (px[15] << 16)) : "memory");
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------ REMAP warp-tile
// predicated 64-bit store: keeps the unrolled row loop one basic block (no BSSY/BRA around the stores)
__device__ __forceinline__ void stg_v2_if(void *p, uint32_t a, uint32_t b, bool on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};\n\t}" ::"l"(p), "r"(a),
                 "r"(b), "r"((uint32_t)on)
                 : "memory");
}

template <int MODE, int DM, bool SWAP, bool TILED, int PB = 0>
__device__ __forceinline__ void remap_tile(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                           const double (&b)[8])
{
    const int lane = C.lane, ns = C.ns;
    const int n_chunks = (T.n_rows + 3 + RC - 1) / RC;
    const int x0 = T.src_x0 & ~7; // box origin; window column 0 sits (src_x0 & 7) samples in
    const uint32_t offL = 2u * (uint32_t)((T.src_x0 - x0) & ~3) + 8u * (uint32_t)lane;
    const uint32_t offR = 2u * (uint32_t)((T.src_x0 - x0 + T.half) & ~3) + 8u * (uint32_t)lane;
    {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage<TILED, PB>(P, T, C, c, x0, T.src_y0 + c * RC);
    }
    // 2-D weights w[r][c] = fl32(wy[r] * wx[c]), identical for the whole tile
    const f2 nz = *reinterpret_cast<const f2 *>(P.tab + 128);
    f2 W[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float wv = __fmul_rn(__ldg(P.tab + 4 * T.fy + r), __ldg(P.tab + 4 * T.fx + c));
            W[r][c] = pk(wv, wv);
        }
    const bool active = 4 * lane < T.half;
    const int64_t pitch = P.out_pitch;
    uint16_t *oL = P.out + T.out_off + 4 * lane - 3 * pitch; // output row (m - 3) while source row m is consumed
    uint16_t *oR = oL + T.half;
    const int n_rows = T.n_rows;
    f2 B0[4], B1[4], B2[4], B3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) B0[o] = B1[o] = B2[o] = B3[o] = 0ull;

    // swap + RRC + float of this lane's 4 + 4 samples of one staged row, then the 3 + 3 window columns of lane+1
    auto convert = [&](uint32_t sa, int rr, f2(&win)[7]) {
        float fl[4], fr[4];
#if OIP_DBG_VARIANT == 1 || OIP_DBG_VARIANT == 2 // timing experiment: no conversion at all
        {
            const uint2 A = lds64(sa + offL + rr * ROW_BYTES), B = lds64(sa + offR + rr * ROW_BYTES);
            fl[0] = __uint_as_float(A.x); fl[1] = __uint_as_float(A.y); fl[2] = fl[0]; fl[3] = fl[1];
            fr[0] = __uint_as_float(B.x); fr[1] = __uint_as_float(B.y); fr[2] = fr[0]; fr[3] = fr[1];
        }
#else
        convert4<MODE, DM, SWAP>(sa + offL + rr * ROW_BYTES, k, b, fl);
        convert4<MODE, DM, SWAP>(sa + offR + rr * ROW_BYTES, k + 4, b + 4, fr);
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) win[j] = pk(fl[j], fr[j]);
#pragma unroll
        for (int j = 0; j < 3; ++j) win[4 + j] = shfl_down1(win[j]);
    };
    // one source row into the 4 output rows it feeds.  AN: new accumulator (weight row 0), A1..A3 receive weight
    // rows 1..3; A3 completes here: per row ((s0*w0 + s1*w1) + s2*w2) + s3*w3, rows accumulated in order
    // 0,1,2,3 (OpenCV's interior order), no FMA contraction
    auto resample = [&](const f2(&win)[7], int m, f2(&AN)[4], f2(&A1)[4], f2(&A2)[4], f2(&A3)[4]) {
        f2 out[4];
#if OIP_DBG_VARIANT == 3 // timing experiment: no FP32 work
#pragma unroll
        for (int o = 0; o < 4; ++o) { out[o] = win[o] ^ win[o + 3] ^ AN[o]; AN[o] = A1[o]; }
#else
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            auto dot = [&](const f2(&Wr)[4]) {
                return add2(add2(add2(mul2(win[o], Wr[0], nz), mul2(win[o + 1], Wr[1], nz)), mul2(win[o + 2], Wr[2], nz)),
                            mul2(win[o + 3], Wr[3], nz));
            };
            AN[o] = dot(W[0]);
            A1[o] = add2(A1[o], dot(W[1]));
            A2[o] = add2(A2[o], dot(W[2]));
            out[o] = add2(A3[o], dot(W[3]));
        }
#endif
        const bool on = active && (unsigned)(m - 3) < (unsigned)n_rows;
#if OIP_DBG_VARIANT == 2 // timing experiment: no F2I / pack, one store
        stg_v2_if(oL, (uint32_t)(out[0] ^ out[1]), (uint32_t)((out[2] ^ out[3]) >> 32), on);
#else
        const uint2 vl = make_uint2(pack16(cast_u16(lo_of(out[0])), cast_u16(lo_of(out[1]))), pack16(cast_u16(lo_of(out[2])), cast_u16(lo_of(out[3]))));
        const uint2 vr = make_uint2(pack16(cast_u16(hi_of(out[0])), cast_u16(hi_of(out[1]))), pack16(cast_u16(hi_of(out[2])), cast_u16(hi_of(out[3]))));
        if (on) { // two predicated stores (streaming: written once, never re-read)
            __stcs(reinterpret_cast<uint2 *>(oL), vl);
            __stcs(reinterpret_cast<uint2 *>(oR), vr);
        }
#endif
        oL += pitch;
        oR += pitch;
    };

    // software pipeline: the conversion of source row m+1 (long XU / FP64 / SHFL latency chain) is issued ahead
    // of the FP32 work of row m, so the two overlap inside one warp
    int slot = 0;
    uint32_t phase = 0;
    mbar_wait_u32(C.bar0, 0);
    uint32_t sa = C.stage0;
    if (PB) { unpack_stage<PB>(C, sa, x0); sa = C.le0; } // packed: rows are converted from the unpacked image of the stage
    f2 wa[7], wb[7];
    convert(sa, 0, wa);
    for (int c = 0; c + 1 < n_chunks; ++c) {
        const int m = c * RC;
        convert(sa, 1, wb);
        resample(wa, m, B0, B1, B2, B3);
        convert(sa, 2, wa);
        resample(wb, m + 1, B3, B0, B1, B2);
        convert(sa, 3, wb);
        resample(wa, m + 2, B2, B3, B0, B1);
        // stage `slot` is consumed: refill it, move on to the next stage and convert its first row
        __syncwarp();
        if (c + ns < n_chunks) issue_stage<TILED, PB>(P, T, C, slot, x0, T.src_y0 + (c + ns) * RC);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        sa = C.stage0 + (uint32_t)slot * STAGE_BYTES;
        if (PB) { // the unpacked image of the previous stage has been consumed (the __syncwarp above ordered its last reads)
            unpack_stage<PB>(C, sa, x0);
            sa = C.le0;
        }
        convert(sa, 0, wa);
        resample(wb, m + 3, B1, B2, B3, B0);
    }
    { // last stage (peeled: the loop body above has no conditional conversion, so its registers line up)
        const int m = (n_chunks - 1) * RC;
        convert(sa, 1, wb);
        resample(wa, m, B0, B1, B2, B3);
        convert(sa, 2, wa);
        resample(wb, m + 1, B3, B0, B1, B2);
        convert(sa, 3, wb);
        resample(wa, m + 2, B2, B3, B0, B1);
        resample(wb, m + 3, B1, B2, B3, B0);
    }
}

// ------------------------------------------------------------------------------------------- COPY warp-tile
template <int MODE, bool SWAP, bool TAIL, bool TILED, int PB = 0>
__device__ __forceinline__ void copy_tile(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                          const double (&b)[8])
{
    const int lane = C.lane, ns = C.ns;
    const int n_rows = T.n_rows;
    const int n_chunks = (n_rows + RC - 1) / RC;
    {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage<TILED, PB>(P, T, C, c, T.x_begin, T.src_y0 + c * RC);
    }
    const bool full = 8 * lane + 8 <= T.half;       // T.half = columns of this strip: any number <= 256
    const int tail = full ? 0 : max(0, T.half - 8 * lane); // the lane that holds the strip's last odd columns
    const int64_t pitch = P.out_pitch;
    uint16_t *o = P.out + T.out_off + 8 * lane;
    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        uint32_t sa = C.stage0 + (uint32_t)slot * STAGE_BYTES + 16u * (uint32_t)lane;
        if (PB) { unpack_stage<PB>(C, C.stage0 + (uint32_t)slot * STAGE_BYTES, T.x_begin); sa = C.le0 + 16u * (uint32_t)lane; }
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const uint4 v = lds128(sa + rr * ROW_BYTES);
            uint32_t wd[4] = {v.x, v.y, v.z, v.w};
            if (SWAP) {
#pragma unroll
                for (int i = 0; i < 4; ++i) wd[i] = __byte_perm(wd[i], 0u, 0x2301);
            }
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const D2 d = split_word(wd[i]);
                    wd[i] = pack16(rrc_d<MODE>(d.lo, k[2 * i], b[2 * i]), rrc_d<MODE>(d.hi, k[2 * i + 1], b[2 * i + 1]));
                }
            }
            if (c * RC + rr < n_rows) {
                if (full) {
                    stg_na_v4(o, make_uint4(wd[0], wd[1], wd[2], wd[3]));
                } else if (TAIL) { // the strip's last odd columns (planner: only the last strip of a CCD can have them)
#pragma unroll
                    for (int j = 0; j < 7; ++j)
                        if (j < tail) o[j] = (uint16_t)(wd[j >> 1] >> (16 * (j & 1)));
                }
            }
            o += pitch;
        }
        __syncwarp();
        if (c + ns < n_chunks) issue_stage<TILED, PB>(P, T, C, slot, T.x_begin, T.src_y0 + (c + ns) * RC);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------- EDGE warp-tile
// Column slivers next to the REMAP spans: image-border columns (4x4 footprint partly outside the CCD) and the
// few columns the 8-column span alignment leaves over, over the same regular interior ROW runs.  One output
// column per lane, scalar FP32, both of OpenCV's accumulation orders (interior: per-row sums added row by row;
// border: one flat left-to-right chain starting from 0) computed and selected per column.  Taps
// outside the CCD are zero: TMA zero-fills them and their (k,b) are forced to 0.  < 0.1 % of the pixels.
template <bool TILED, int PB = 0>
__device__ __forceinline__ void edge_tile(const FastParams &P, const FastTile &T, const WarpCtx &C)
{
    const int lane = C.lane, ns = C.ns;
    const int n_chunks = (T.n_rows + 3 + RC - 1) / RC;
    const int x0 = T.src_x0 & ~7;                  // floor to a multiple of 8 (also for negative columns)
    const int col = T.src_x0 + lane;               // source column this lane converts = first tap of output column x_begin+lane
    const uint32_t off = 2u * (uint32_t)(col - x0);
    {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage<TILED, PB>(P, T, C, c, x0, T.src_y0 + c * RC);
    }
    const bool swap = P.ccd[T.ccd].swap != 0;
    const double *kbp = P.ccd[T.ccd].kb;
    const bool inside = col >= 0 && col < P.w;
    double k = inside ? 1.0 : 0.0, b = 0.0;
    if (kbp && inside) { k = kbp[2 * (int64_t)col]; b = kbp[2 * (int64_t)col + 1]; }
    float W[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) W[r][c] = __fmul_rn(__ldg(P.tab + 4 * T.fy + r), __ldg(P.tab + 4 * T.fx + c));
    const bool active = lane < T.half;
    const bool border = !(col >= 0 && col < P.w - 3); // the footprint of my output column leaves the CCD: flat order
    uint16_t *o = P.out + T.out_off + lane - 3 * P.out_pitch;
    float A[4] = {0.f, 0.f, 0.f, 0.f}, F[4] = {0.f, 0.f, 0.f, 0.f}; // partial sums of the output rows in flight (index 0 unused)
    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        uint32_t sa = C.stage0 + (uint32_t)slot * STAGE_BYTES + off;
        if (PB) { unpack_stage<PB>(C, C.stage0 + (uint32_t)slot * STAGE_BYTES, x0); sa = C.le0 + off; }
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            uint32_t s;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(s) : "r"(sa + rr * ROW_BYTES));
            if (swap) s = __byte_perm(s, 0u, 0x4401);
            const float v = inside ? (float)(kbp ? rrc_px(s, k, b) : s) : 0.f;
            float win[4];
            win[0] = v;
#pragma unroll
            for (int j = 1; j < 4; ++j) win[j] = __shfl_down_sync(0xffffffffu, v, j);
            // A[i] / F[i]: interior / flat partial sum of the output row that has received weight rows 0..i-1 so far
            float d[4], f[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float p0 = __fmul_rn(win[0], W[r][0]), p1 = __fmul_rn(win[1], W[r][1]), p2 = __fmul_rn(win[2], W[r][2]),
                            p3 = __fmul_rn(win[3], W[r][3]);
                d[r] = __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3);
                f[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(r == 0 ? 0.f : F[r], p0), p1), p2), p3);
            }
            const float a_out = __fadd_rn(A[3], d[3]), f_out = f[3];
            A[3] = __fadd_rn(A[2], d[2]); F[3] = f[2];
            A[2] = __fadd_rn(A[1], d[1]); F[2] = f[1];
            A[1] = d[0];                  F[1] = f[0];
            const int m = c * RC + rr;
            if (active && (unsigned)(m - 3) < (unsigned)T.n_rows) *o = (uint16_t)cast_u16(border ? f_out : a_out);
            o += P.out_pitch;
        }
        __syncwarp();
        if (c + ns < n_chunks) issue_stage<TILED, PB>(P, T, C, slot, x0, T.src_y0 + (c + ns) * RC);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
    }
}

template <int MODE, bool SWAP, bool TILED, int PB = 0>
__device__ __forceinline__ void remap_dispatch(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                               const double (&b)[8])
{
    const int dm = T.src_x0 & 3;
    if (dm == 0) remap_tile<MODE, 0, SWAP, TILED, PB>(P, T, C, k, b);
    else if (dm == 1) remap_tile<MODE, 1, SWAP, TILED, PB>(P, T, C, k, b);
    else if (dm == 2) remap_tile<MODE, 2, SWAP, TILED, PB>(P, T, C, k, b);
    else remap_tile<MODE, 3, SWAP, TILED, PB>(P, T, C, k, b);
}

template <int MODE, int PB>
__device__ __forceinline__ void packed_dispatch(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                                const double (&b)[8])
{
    if (T.kind == FT_REMAP) remap_dispatch<MODE, false, false, PB>(P, T, C, k, b); // the unpacked image holds native u16
    else if (T.half & 7) copy_tile<MODE, false, true, false, PB>(P, T, C, k, b);
    else copy_tile<MODE, false, false, false, PB>(P, T, C, k, b);
}

template <int MODE>
__device__ __forceinline__ void tile_dispatch(const FastParams &P, const FastTile &T, const WarpCtx &C, bool swap, bool tiled,
                                              const double (&k)[8], const double (&b)[8])
{
    const int pb = P.ccd[T.ccd].pbits;
    if (pb == 12) { packed_dispatch<MODE, 12>(P, T, C, k, b); return; }
    if (pb == 10) { packed_dispatch<MODE, 10>(P, T, C, k, b); return; }
    if (T.kind == FT_REMAP) {
        if (tiled) remap_dispatch<MODE, true, true>(P, T, C, k, b); // sub-images hold big-endian samples
        else if (swap) remap_dispatch<MODE, true, false>(P, T, C, k, b);
        else remap_dispatch<MODE, false, false>(P, T, C, k, b);
    } else {
        if (tiled) {
            if (T.half & 7) copy_tile<MODE, true, true, true>(P, T, C, k, b);
            else copy_tile<MODE, true, false, true>(P, T, C, k, b);
        } else if (T.half & 7) { // rare: keep the tail stores out of the common loop
            if (swap) copy_tile<MODE, true, true, false>(P, T, C, k, b);
            else copy_tile<MODE, false, true, false>(P, T, C, k, b);
        } else {
            if (swap) copy_tile<MODE, true, false, false>(P, T, C, k, b);
            else copy_tile<MODE, false, false, false>(P, T, C, k, b);
        }
    }
}

// one warp-tile, start to finish (the warp's stage ring and barriers are idle on entry and on return)
__device__ __forceinline__ void run_tile(const FastParams &P, const FastTile &T, WarpCtx &C)
{
    const int lane = C.lane;
    const bool tiled = P.ccd[T.ccd].tiled != 0;
    C.tm = tiled ? nullptr : &P.tmap[T.tmap];
    if (lane == 0) {
        // a stage completes on the tensor copy's byte count (one arrival), or on the 32 lanes' cp.async completions
        for (int s = 0; s < C.ns; ++s) mbar_init_u32(C.bar0 + 8u * s, tiled ? 32u : 1u);
        fence_mbar_init();
    }
    __syncwarp();
    if (T.kind == FT_EDGE) {
        const int pb = P.ccd[T.ccd].pbits;
        if (tiled) edge_tile<true>(P, T, C);
        else if (pb == 12) edge_tile<false, 12>(P, T, C);
        else if (pb == 10) edge_tile<false, 10>(P, T, C);
        else edge_tile<false>(P, T, C);
        return;
    }
    // (k,b) of the 8 detectors this lane converts, and the warp-wide RRC mode
    const double *kbp = P.ccd[T.ccd].kb;
    double k[8], b[8];
    bool general = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int col = T.kind == FT_REMAP ? T.src_x0 + 4 * lane + (j & 3) + (j >> 2) * T.half : T.x_begin + 8 * lane + j;
        col = min(col, P.w - 1); // lanes past the tile's window convert (unused) duplicates of the last detector
        k[j] = 1.0;
        b[j] = 0.0;
        if (kbp) {
            const double2 v = *reinterpret_cast<const double2 *>(kbp + 2 * (int64_t)col);
            k[j] = v.x;
            b[j] = v.y;
            general = general || !(v.x >= 0.0 && v.y >= 0.0 && __dadd_rn(__dmul_rn(v.x, 65535.0), v.y) < 2147483648.0);
        }
    }
    const bool swap = P.ccd[T.ccd].swap != 0;
    const int mode = kbp ? (__any_sync(0xffffffffu, general) ? 2 : 1) : 0;
    if (mode == 1) tile_dispatch<1>(P, T, C, swap, tiled, k, b);
    else if (mode == 0) tile_dispatch<0>(P, T, C, swap, tiled, k, b);
    else tile_dispatch<2>(P, T, C, swap, tiled, k, b);
}

// MINB = CTAs per SM the register allocation must allow (4: 128 registers, 3: 168)
template <int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) pan_fast_kernel(const __grid_constant__ FastParams P)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[WARPS][MAX_STAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpCtx C;
    C.ns = P.n_stage;
    C.lane = lane;
    C.tm = nullptr;
    C.stage0 = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(warp * C.ns) * STAGE_BYTES;
    C.le0 = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(WARPS * C.ns + warp) * STAGE_BYTES; // (allocated when a CCD is packed)
    C.bar0 = smem_u32(&bars[warp][0]);
    const FastTile T = P.tiles[(int64_t)blockIdx.x * WARPS + warp];
    if (T.kind < 0) return; // padding entry of a partial CTA; warps never synchronise with each other
    run_tile(P, T, C);
}

// ------------------------------------------------------------------------------------------------ host side
} // namespace panfast

namespace tmaw {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int encode_tmap_u32(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes, int box_w32, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)(w / 2), (cuuint64_t)n_rows}; // 32-bit elements = sample pairs
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return OIP_OK;
}
} // namespace tmaw

namespace panfast {
int fast_encode_tmap(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes)
{
    return tmaw::encode_tmap_u32(tm, base, w, n_rows, pitch_bytes, BOX_W, RC);
}

int fast_encode_tmap_packed(CUtensorMap *tm, const void *base, int w, int bits, int64_t n_rows, int64_t pitch_bytes)
{
    // the packed line as 32-bit elements (w * bits / 32 of them): encode_tmap_u32 counts w / 2 elements for w samples
    return tmaw::encode_tmap_u32(tm, base, 2 * (w * bits / 32), n_rows, pitch_bytes, bits == 12 ? PBOX12 : PBOX10, RC);
}

int fast_launch(oip_ctx *ctx, const FastParams &P, int64_t n_ctas)
{
    bool packed = false;
    for (int i = 0; i < 8; ++i) packed = packed || P.ccd[i].pbits != 0;
    const size_t smem = (size_t)WARPS * (P.n_stage + (packed ? 1 : 0)) * STAGE_BYTES + 128;
    if (!ctx->fast_attr_set) {
        OIP_CUDA(cudaFuncSetAttribute(pan_fast_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * (MAX_STAGE + 1) * STAGE_BYTES + 128));
        ctx->fast_attr_set = true;
    }
    // (a 4-CTAs/SM, 128-register instantiation was measured in round 1: no faster, heavy spills -- dropped)
    pan_fast_kernel<3><<<(unsigned)n_ctas, WARPS * 32, smem, ctx->stream>>>(P);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

} // namespace panfast
} // namespace oip
