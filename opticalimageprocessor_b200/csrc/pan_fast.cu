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

namespace oip {
namespace panfast {

constexpr int ROW_BYTES = BOX_W * 2;                          // 272
constexpr int BOX_BYTES = ROW_BYTES * RC;                     // 1088
constexpr int BOX_STRIDE = (BOX_BYTES + 127) / 128 * 128;     // 1152: TMA destinations are 128-byte aligned
constexpr int STAGE_BYTES = 2 * BOX_STRIDE;                   // two boxes per stage
constexpr int MAX_STAGE = 8;
static_assert(RC == 4, "the row loop is unrolled by the 4-deep accumulator rotation");

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void stg_v2(void *p, uint32_t a, uint32_t b)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// cvRound + saturate_cast<ushort>: PTX float->int conversions clamp to the destination range (SASS F2I.U16.NTZ)
__device__ __forceinline__ uint32_t cast_u16(float s)
{
    unsigned short r;
    asm("cvt.rni.u16.f32 %0, %1;" : "=h"(r) : "f"(s));
    return r;
}
__device__ __forceinline__ f2 shfl_down1(f2 v)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_down_sync(0xffffffffu, lo, 1);
    hi = __shfl_down_sync(0xffffffffu, hi, 1);
    return ((f2)hi << 32) | lo;
}

// RRC modes: 0 = none (stitch only), 1 = every (k,b) of the warp is >= 0 and k*65535+b < 2^31 (exact with two
// magic adds, no range test per pixel), 2 = general (sign / range handling exactly like x86 cvttsd2si)
template <int MODE>
__device__ __forceinline__ uint32_t rrc_mode(uint32_t s, double k, double b)
{
    if (MODE == 0) return s;
    if (MODE == 1) {
        const double sd = __dadd_rn(__hiloint2double(0x43300000, (int)s), -4503599627370496.0);
        const double v = __dadd_rn(__dmul_rn(k, sd), b);
        return (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0)); // low 16 bits are taken by the caller
    }
    return rrc_px(s, k, b);
}

struct WarpCtx {
    const CUtensorMap *tm;
    uint32_t stage0, bar0; // shared-memory addresses of this warp's stage ring and barriers
    int ns, lane;
    uint32_t sel_lo, sel_hi; // PRMT selectors: halfword -> zero-extended (byte-swapped) sample
};

__device__ __forceinline__ void issue_stage(const WarpCtx &C, int slot, int xa, int xb, int y)
{
    const uint32_t bar = C.bar0 + 8u * slot, dst = C.stage0 + (uint32_t)slot * STAGE_BYTES;
    mbar_expect_tx_u32(bar, 2 * BOX_BYTES);
    tma_load_2d(dst, C.tm, xa, y, bar);
    tma_load_2d(dst + BOX_STRIDE, C.tm, xb, y, bar);
}

// ------------------------------------------------------------------------------------------ REMAP warp-tile
// The TMA unit only accepts box origins on 16-byte boundaries of the tensor row (measured: tools/tma_probe.cu,
// any other x coordinate raises "illegal instruction"), so a box starts at the source window's column rounded
// down to a multiple of 8 and the lanes read their 4 samples DM = (window column mod 4) halfwords into an
// aligned 8-byte word: DM is a template parameter, the loads and PRMT selections stay fixed.
template <int DM>
__device__ __forceinline__ void load4(uint32_t a, uint32_t sel_lo, uint32_t sel_hi, uint32_t *s)
{
    const uint2 A = lds64(a);
    if (DM == 0) {
        s[0] = __byte_perm(A.x, 0u, sel_lo); s[1] = __byte_perm(A.x, 0u, sel_hi);
        s[2] = __byte_perm(A.y, 0u, sel_lo); s[3] = __byte_perm(A.y, 0u, sel_hi);
    } else if (DM == 1) {
        const uint32_t B = lds32(a + 8);
        s[0] = __byte_perm(A.x, 0u, sel_hi); s[1] = __byte_perm(A.y, 0u, sel_lo);
        s[2] = __byte_perm(A.y, 0u, sel_hi); s[3] = __byte_perm(B, 0u, sel_lo);
    } else if (DM == 2) {
        const uint32_t B = lds32(a + 8);
        s[0] = __byte_perm(A.y, 0u, sel_lo); s[1] = __byte_perm(A.y, 0u, sel_hi);
        s[2] = __byte_perm(B, 0u, sel_lo); s[3] = __byte_perm(B, 0u, sel_hi);
    } else {
        const uint2 B = lds64(a + 8);
        s[0] = __byte_perm(A.y, 0u, sel_hi); s[1] = __byte_perm(B.x, 0u, sel_lo);
        s[2] = __byte_perm(B.x, 0u, sel_hi); s[3] = __byte_perm(B.y, 0u, sel_lo);
    }
}

template <int MODE, int DM>
__device__ __forceinline__ void remap_tile(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                           const double (&b)[8])
{
    const int lane = C.lane, ns = C.ns;
    const int n_chunks = (T.n_rows + 3 + RC - 1) / RC;
    const int xa = T.src_x0 & ~7, xb = (T.src_x0 + T.half) & ~7; // box origins; window column 0 sits (src_x0 & 7) samples in
    const uint32_t offL = 8u * (uint32_t)(((T.src_x0 & 7) >> 2) + lane);
    const uint32_t offR = BOX_STRIDE + 8u * (uint32_t)((((T.src_x0 + T.half) & 7) >> 2) + lane);
    if (lane == 0) {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage(C, c, xa, xb, T.src_y0 + c * RC);
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
    const int half = T.half, n_rows = T.n_rows;
    f2 B0[4], B1[4], B2[4], B3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) B0[o] = B1[o] = B2[o] = B3[o] = 0ull;

    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        const uint32_t sa = C.stage0 + (uint32_t)slot * STAGE_BYTES;
        // AN: new accumulator (weight row 0), A1..A3: rows that receive weight rows 1..3; A3 completes here
        auto row = [&](int rr, f2(&AN)[4], f2(&A1)[4], f2(&A2)[4], f2(&A3)[4]) {
            uint32_t s[8];
            load4<DM>(sa + offL + rr * ROW_BYTES, C.sel_lo, C.sel_hi, s);
            load4<DM>(sa + offR + rr * ROW_BYTES, C.sel_lo, C.sel_hi, s + 4);
            f2 win[7];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float fl = (float)(uint16_t)rrc_mode<MODE>(s[j], k[j], b[j]);
                const float fr = (float)(uint16_t)rrc_mode<MODE>(s[4 + j], k[4 + j], b[4 + j]);
                win[j] = pk(fl, fr);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) win[4 + j] = shfl_down1(win[j]);
            f2 out[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                // per row ((s0*w0 + s1*w1) + s2*w2) + s3*w3, rows accumulated in order 0,1,2,3 (OpenCV interior order)
                auto dot = [&](const f2(&Wr)[4]) {
                    return add2(add2(add2(mul2(win[o], Wr[0], nz), mul2(win[o + 1], Wr[1], nz)), mul2(win[o + 2], Wr[2], nz)),
                                mul2(win[o + 3], Wr[3], nz));
                };
                AN[o] = dot(W[0]);
                A1[o] = add2(A1[o], dot(W[1]));
                A2[o] = add2(A2[o], dot(W[2]));
                out[o] = add2(A3[o], dot(W[3]));
            }
            const int m = c * RC + rr;
            if (active && (unsigned)(m - 3) < (unsigned)n_rows) {
                stg_v2(oL, cast_u16(lo_of(out[0])) | (cast_u16(lo_of(out[1])) << 16),
                       cast_u16(lo_of(out[2])) | (cast_u16(lo_of(out[3])) << 16));
                stg_v2(oL + half, cast_u16(hi_of(out[0])) | (cast_u16(hi_of(out[1])) << 16),
                       cast_u16(hi_of(out[2])) | (cast_u16(hi_of(out[3])) << 16));
            }
            oL += pitch;
        };
        row(0, B0, B1, B2, B3);
        row(1, B3, B0, B1, B2);
        row(2, B2, B3, B0, B1);
        row(3, B1, B2, B3, B0);
        __syncwarp();
        if (lane == 0 && c + ns < n_chunks) issue_stage(C, slot, xa, xb, T.src_y0 + (c + ns) * RC);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------- COPY warp-tile
template <int MODE>
__device__ __forceinline__ void copy_tile(const FastParams &P, const FastTile &T, const WarpCtx &C, const double (&k)[8],
                                          const double (&b)[8])
{
    const int lane = C.lane, ns = C.ns;
    const int n_rows = T.n_rows;
    const int n_chunks = (n_rows + RC - 1) / RC;
    const int xa = T.x_begin, xb = T.x_begin + 128;
    if (lane == 0) {
        const int pre = min(ns, n_chunks);
        for (int c = 0; c < pre; ++c) issue_stage(C, c, xa, xb, T.src_y0 + c * RC);
    }
    const bool active = 8 * lane < T.half;
    const int64_t pitch = P.out_pitch;
    uint16_t *o = P.out + T.out_off + 8 * lane;
    const uint32_t my = (uint32_t)(lane >> 4) * BOX_STRIDE + (uint32_t)(lane & 15) * 16u;
    int slot = 0;
    uint32_t phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait_u32(C.bar0 + 8u * slot, phase);
        const uint32_t sa = C.stage0 + (uint32_t)slot * STAGE_BYTES + my;
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const uint4 v = lds128(sa + rr * ROW_BYTES);
            const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
            uint32_t t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                t[j] = rrc_mode<MODE>(__byte_perm(wd[j >> 1], 0u, (j & 1) ? C.sel_hi : C.sel_lo), k[j], b[j]);
            if (active && c * RC + rr < n_rows) {
                uint4 w;
                w.x = __byte_perm(t[0], t[1], 0x5410);
                w.y = __byte_perm(t[2], t[3], 0x5410);
                w.z = __byte_perm(t[4], t[5], 0x5410);
                w.w = __byte_perm(t[6], t[7], 0x5410);
                stg_na_v4(o, w);
            }
            o += pitch;
        }
        __syncwarp();
        if (lane == 0 && c + ns < n_chunks) issue_stage(C, slot, xa, xb, T.src_y0 + (c + ns) * RC);
        if (++slot == ns) { slot = 0; phase ^= 1u; }
    }
}

__global__ void __launch_bounds__(WARPS * 32, 4) pan_fast_kernel(const __grid_constant__ FastParams P)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[WARPS][MAX_STAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FastTile T = P.tiles[(int64_t)blockIdx.x * WARPS + warp];
    if (T.kind < 0) return; // padding entry of a partial CTA; warps never synchronise with each other

    WarpCtx C;
    C.ns = P.n_stage;
    C.lane = lane;
    C.tm = &P.tmap[T.tmap];
    C.stage0 = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(warp * C.ns) * STAGE_BYTES;
    C.bar0 = smem_u32(&bars[warp][0]);
    const bool swap = P.ccd[T.ccd].swap != 0;
    C.sel_lo = swap ? 0x4401u : 0x4410u;
    C.sel_hi = swap ? 0x4423u : 0x4432u;
    if (lane == 0) {
        for (int s = 0; s < C.ns; ++s) mbar_init_u32(C.bar0 + 8u * s, 1);
        fence_mbar_init();
    }
    __syncwarp();

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
    const int mode = kbp ? (__any_sync(0xffffffffu, general) ? 2 : 1) : 0;
    if (T.kind == FT_REMAP) {
        const int dm = T.src_x0 & 3;
#define OIP_REMAP_DM(M)                                           \
    do {                                                          \
        if (dm == 0) remap_tile<M, 0>(P, T, C, k, b);             \
        else if (dm == 1) remap_tile<M, 1>(P, T, C, k, b);        \
        else if (dm == 2) remap_tile<M, 2>(P, T, C, k, b);        \
        else remap_tile<M, 3>(P, T, C, k, b);                     \
    } while (0)
        if (mode == 1) OIP_REMAP_DM(1);
        else if (mode == 0) OIP_REMAP_DM(0);
        else OIP_REMAP_DM(2);
#undef OIP_REMAP_DM
    } else {
        if (mode == 1) copy_tile<1>(P, T, C, k, b);
        else if (mode == 0) copy_tile<0>(P, T, C, k, b);
        else copy_tile<2>(P, T, C, k, b);
    }
}

// ------------------------------------------------------------------------------------------------ host side
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

int fast_encode_tmap(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {BOX_W, RC};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return OIP_OK;
}

int fast_launch(oip_ctx *ctx, const FastParams &P, int64_t n_ctas)
{
    const size_t smem = (size_t)WARPS * P.n_stage * STAGE_BYTES + 128;
    if (!ctx->fast_attr_set) {
        OIP_CUDA(cudaFuncSetAttribute(pan_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      WARPS * MAX_STAGE * STAGE_BYTES + 128));
        ctx->fast_attr_set = true;
    }
    pan_fast_kernel<<<(unsigned)n_ctas, WARPS * 32, smem, ctx->stream>>>(P);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

} // namespace panfast
} // namespace oip
