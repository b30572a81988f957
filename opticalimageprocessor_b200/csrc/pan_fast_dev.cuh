// pan_fast_dev.cuh -- device code of the fused PAN fast path (see pan_fast.cu for the description); compiled once per
// source-format class by pan_fast_c0.cu / pan_fast_c1.cu / pan_fast_c2.cu.
#pragma once
#include "pan_fast.cuh"
#include "tma_warp.cuh"

#ifndef OIP_FAST_CLS
#error "include from pan_fast_c0.cu / _c1.cu / _c2.cu with OIP_FAST_CLS = 0 (line rasters), 1 (sub-image tiles), 2 (packed lines)"
#endif
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
    // the run-time -0.0 addend of the packed products (mul2()) as ONE register broadcast to both halves: the product is
    // FFMA2 pair, W.F32, Z.F32 -- two vector-register reads fewer than with a (-0,-0) pair
    // (tools/probes/ffma2_rate.cu: 2.7 against 3.6 cycles per instruction and sub-partition; 3.79 -> 3.69 ms on 131072 lines)
    const float nzs = __ldg(P.tab + 128);
    const f2 nz = pk(nzs, nzs);
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
    const int n_rows = T.n_rows;
    // stores: two constant base addresses + ONE 32-bit row offset stepped per row (two stepped 64-bit pointers cost
    // IADD3 + IADD3.X + two MOVs into an aligned pair each: 3.69 -> 3.59 ms); the planner keeps rows x pitch < 2^32
    const uint64_t bL = reinterpret_cast<uint64_t>(P.out + T.out_off + 4 * lane), bR = bL + 2ull * (uint64_t)T.half;
    const uint32_t st_pitch = (uint32_t)(2 * pitch);
    uint32_t st_off = 0u - 3u * st_pitch; // output row (m - 3) while source row m is consumed; rows < 0 are predicated off
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
        stg_v2_if(reinterpret_cast<void *>(bL + st_off), (uint32_t)(out[0] ^ out[1]), (uint32_t)((out[2] ^ out[3]) >> 32), on);
        st_off += st_pitch;
#else
        const uint2 vl = make_uint2(pack16(cast_u16(lo_of(out[0])), cast_u16(lo_of(out[1]))), pack16(cast_u16(lo_of(out[2])), cast_u16(lo_of(out[3]))));
        const uint2 vr = make_uint2(pack16(cast_u16(hi_of(out[0])), cast_u16(hi_of(out[1]))), pack16(cast_u16(hi_of(out[2])), cast_u16(hi_of(out[3]))));
        { // two predicated stores (written once, never re-read: no L1 allocation)
            uint64_t aL, aR;
            asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(aL) : "r"(st_off), "l"(bL));
            asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(aR) : "r"(st_off), "l"(bR));
            stg_v2_if(reinterpret_cast<void *>(aL), vl.x, vl.y, on);
            stg_v2_if(reinterpret_cast<void *>(aR), vr.x, vr.y, on);
            st_off += st_pitch;
        }
#endif
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
    constexpr int CLS = OIP_FAST_CLS;
    if (CLS < 0 || CLS == 2) {
        const int pb = P.ccd[T.ccd].pbits;
        if (pb == 12) { packed_dispatch<MODE, 12>(P, T, C, k, b); return; }
        if (pb == 10 || CLS == 2) { packed_dispatch<MODE, 10>(P, T, C, k, b); return; }
    }
    if (CLS == 1) tiled = true;
    if (CLS == 0) tiled = false;
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
#ifdef OIP_FAST_LEAN
        edge_tile<false>(P, T, C);
        return;
#endif
        constexpr int CLS = OIP_FAST_CLS;
        const int pb = P.ccd[T.ccd].pbits;
        if (CLS == 1 || (CLS < 0 && tiled)) edge_tile<true>(P, T, C);
        else if (CLS == 2 || (CLS < 0 && pb != 0)) {
            if (pb == 12) edge_tile<false, 12>(P, T, C);
            else edge_tile<false, 10>(P, T, C);
        } else edge_tile<false>(P, T, C);
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
#ifdef OIP_FAST_LEAN // kernel experiments only (tools/probes/build_variant.py): big-endian rasters with RRC, nothing else
    if (T.kind == FT_REMAP) remap_dispatch<1, true, false>(P, T, C, k, b);
    else copy_tile<1, true, false, false>(P, T, C, k, b);
    return;
#endif
    if (mode == 1) tile_dispatch<1>(P, T, C, swap, tiled, k, b);
    else if (mode == 0) tile_dispatch<0>(P, T, C, swap, tiled, k, b);
    else tile_dispatch<2>(P, T, C, swap, tiled, k, b);
}

// One kernel per source-format class (template argument = OIP_FAST_CLS of the translation unit): a strip's CCDs are almost
// always of one class, a third of the code per launch runs 1.3 % faster (instruction cache) and the three translation
// units compile in parallel.  3 CTAs per SM = 168 registers (a 4-CTA, 128-register build was measured in round 1: no
// faster, heavy spills).
template <int CLS>
__global__ void __launch_bounds__(WARPS * 32, 3) pan_fast_kernel(const __grid_constant__ FastParams P, int64_t tile0)
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
    const FastTile T = P.tiles[tile0 + (int64_t)blockIdx.x * WARPS + warp];
    if (T.kind < 0) return; // padding entry of a partial CTA; warps never synchronise with each other
    run_tile(P, T, C);
}


template <>
int fast_launch_cls<OIP_FAST_CLS>(oip_ctx *ctx, const FastParams &P, int64_t tile0, int64_t n_ctas)
{
    constexpr int CLS = OIP_FAST_CLS;
    const size_t smem = (size_t)WARPS * (P.n_stage + (CLS == 2 ? 1 : 0)) * STAGE_BYTES + 128; // packed: + the unpacked image of a stage
    if (!ctx->fast_attr_set[CLS]) {
        OIP_CUDA(cudaFuncSetAttribute(pan_fast_kernel<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * (MAX_STAGE + 1) * STAGE_BYTES + 128));
        ctx->fast_attr_set[CLS] = true;
    }
    pan_fast_kernel<CLS><<<(unsigned)n_ctas, WARPS * 32, smem, ctx->stream>>>(P, tile0);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

} // namespace panfast
} // namespace oip
