// pan_pipeline.cu -- fused unpack -> RRC -> sectioned cubic shift -> trimmed concat (PAN strips).
//
// One launch covers every CCD of a strip.  Work is cut into tiles of TW output columns x up to TH
// output rows; a CTA marches down its tile in chunks of RC rows:
//
//   load     TMA-unit bulk copies (cp.async.bulk, SASS UBLKCP) stage the raw source rows of the next
//            two chunks in shared memory behind an mbarrier while the current chunk is processed:
//            every source byte is read from HBM once per column strip (halo columns hit L2)
//   convert  byte swap (one PRMT per sample) + fp64 RRC with (k,b) held in registers
//            (ref imageop.h:134); results go as floats into a 40-row ring laid out as LANE PAIRS:
//            ring slot s of a row holds (column s, column s+HALF) of the tile's source window
//   resample a thread owns 2 slots x 8 rows = 32 pixels; 128-bit shared loads deliver ready-made
//            packed operands, so OpenCV's bicubic sum (its own order, no FMA) runs as
//            FMUL2/FADD2 on pixel pairs (left half, right half) with no register shuffling at all
//
// The kernel is instruction-issue bound, not HBM bound (profiles/): everything here is about
// instructions per pixel.  No intermediate (.RRC.RAW, .PRESTT.RAW) ever exists in HBM.  CCDs that
// are not shifted use COPY tiles: same staging, RRC, 128-bit stores to their trimmed position.
//
// Irregular situations (map rounding anomalies, section/image borders, partial chunks, unaligned or
// packed inputs) take exact but slower generic paths inside the same kernel.
#include "oip_common.cuh"
#include "pan_fast.cuh"
#include "pan_plan.hpp"

namespace oip {
namespace pan {

#ifndef OIP_PAN_NT
#define OIP_PAN_NT 128
#endif
constexpr int NT = OIP_PAN_NT;         // threads per CTA; everything below derives from it
constexpr int CTAS_PER_SM = NT == 128 ? 4 : 2;
constexpr int SLOTS = NT / 2;          // ring slots per row; slot s = (window col s, window col s+HALF)
constexpr int HALF = SLOTS - 4;        // lane 0 = tile columns [0,HALF), lane 1 = [HALF,TW); +3 taps +1 (map anomaly)
constexpr int TW = 2 * HALF;           // output columns per tile
constexpr int SRC_W = TW + 4;          // source columns of the window
constexpr int SWC = (SRC_W + 7 + 7) / 8 * 8; // staged columns: + up to 7 columns of 16-byte alignment slack
constexpr int RC = 32;                 // output rows per chunk
constexpr int RING = 40;               // ring rows (>= RC + 3 + 1)
constexpr int RING_MIRROR = 10;        // ring rows 0..9 are mirrored behind row RING-1: a thread's 11-row run never wraps
constexpr int RING_ROWS = RING + RING_MIRROR;
constexpr int STG = 40;                // raw staging rows per buffer (first chunk needs RC + 4)
constexpr int TH = 512;                // output rows per tile
constexpr int CV_Q = SLOTS / 4;        // convert step: slot quads per row (power of two)
constexpr int CV_PH = NT / CV_Q;       // convert step: row phases (8)
constexpr int RS_CG = SLOTS / 2;       // resample step: threads per row group (HALF/2 of them active)
constexpr int NWARPS = NT / 32;
static_assert((CV_Q & (CV_Q - 1)) == 0 && NT / RS_CG == 4 && CV_PH == 8, "thread mapping");

enum { KIND_COPY = 0, KIND_REMAP = 1 };

struct Tile {
    int32_t ccd, kind;
    int32_t x_begin, x_end; // CCD columns [x_begin,x_end) produced by this tile
    int32_t out_x;          // output raster column of x_begin
    int32_t n_rows;
    int64_t g0;             // first output row (global)
    int64_t j0;             // section-local row of g0
    int64_t sec_off;        // global source row of buffer row 0
    int64_t stale_off;      // global source row of buffer row 0 for rows >= rows_s, -1 none
    int32_t rows_s, hbuf;
};

struct SegDev {
    const uint8_t *base;
    int64_t row0, n_rows, pitch;
};
struct CcdDev {
    SegDev seg[OIP_MAX_SEG];
    const double *kb;
    double dX, dY;
    const int64_t *tile_off;
    int32_t fmt, n_seg, tile_cols, tile_lines;
};
struct Params {
    CcdDev ccd[8];
    const Tile *tiles;
    uint16_t *out;
    int64_t out_pitch, out_row0;
    const float *tab; // 32x4 cubic weights, then the run-time (-0.0,-0.0) pair
    int *err;
    int32_t w, n_ccd, bulk_ok;
};

__device__ __forceinline__ int dev_sat_short(int v) { return max(-32768, min(32767, v)); }

// cvRound(float(i + d) * 32): ref stitcher.h:96-97 + OpenCV remap fixed-point conversion
__device__ __forceinline__ int dev_map_fixed(int64_t i, double d)
{
    float m = __double2float_rn(__dadd_rn((double)i, d));
    return __float2int_rn(__fmul_rn(m, 32.0f));
}
__device__ __forceinline__ int dev_tap_base(int64_t i, double d) { return dev_sat_short(dev_map_fixed(i, d) >> 5) - 1; }

__device__ __forceinline__ const uint8_t *row_ptr(const CcdDev &C, int64_t g)
{
#pragma unroll
    for (int s = 0; s < OIP_MAX_SEG; ++s)
        if (s < C.n_seg && g >= C.seg[s].row0 && g < C.seg[s].row0 + C.seg[s].n_rows)
            return C.seg[s].base + (g - C.seg[s].row0) * C.seg[s].pitch;
    return nullptr;
}

// buffer row t of the section -> global source row, -1 = zero border
__device__ __forceinline__ int64_t local_to_global(const Tile &T, int64_t t)
{
    if (T.kind == KIND_COPY) return t;
    if (t < 0 || t >= T.hbuf) return -1;
    if (t < T.rows_s) return T.sec_off + t;
    return T.stale_off >= 0 ? T.stale_off + t : -1;
}

// one raw sample for the generic (non-bulk) loader, returned in native byte order
__device__ __forceinline__ uint32_t load_sample(const CcdDev &C, const uint8_t *row, int64_t g, int c)
{
    switch (C.fmt) {
    case OIP_FMT_LE16: return *reinterpret_cast<const uint16_t *>(row + 2 * (int64_t)c);
    case OIP_FMT_BE16: {
        uint32_t v = *reinterpret_cast<const uint16_t *>(row + 2 * (int64_t)c);
        return ((v & 0xFF) << 8) | (v >> 8);
    }
    case OIP_FMT_PACK12: {
        const uint8_t *p = row + (int64_t)(c >> 1) * 3;
        return (c & 1) ? (((uint32_t)(p[1] & 0x0F) << 8) | p[2]) : (((uint32_t)p[0] << 4) | (p[1] >> 4));
    }
    case OIP_FMT_PACK10: {
        const uint8_t *p = row + (int64_t)(c >> 2) * 5;
        int k = c & 3;
        uint32_t hi = p[k], lo = p[k + 1];
        return ((hi << (2 + 2 * k)) | (lo >> (6 - 2 * k))) & 0x3FF;
    }
    default: { // OIP_FMT_BE16_TILES: row == IMDT base, g = global PAN line (ref aux_separator.h:341-372)
        int lpf = 4 * C.tile_lines;
        int64_t f = g / lpf;
        int rl = (int)(g - f * lpf);
        int r = rl / C.tile_lines, y = rl - r * C.tile_lines;
        int cc = c / C.tile_cols, x = c - cc * C.tile_cols;
        int64_t off = C.tile_off[f * 40 + r * 8 + cc];
        if (off < 0) return 0;
        const uint8_t *p = row + off + ((int64_t)y * C.tile_cols + x) * 2;
        return ((uint32_t)p[0] << 8) | p[1];
    }
    }
}

struct ChunkRows {
    int t_lo, t_hi; // source (buffer-local) rows needed by the chunk, inclusive
    int new_lo;     // first row not yet in the ring
    int n_new;
};

__device__ __forceinline__ ChunkRows chunk_rows(const Tile &T, double dY, int k)
{
    ChunkRows c;
    const int64_t ja = T.j0 + (int64_t)k * RC;
    const int64_t jb = min(ja + RC, T.j0 + (int64_t)T.n_rows) - 1;
    if (T.kind == KIND_REMAP) {
        c.t_lo = dev_tap_base(ja, dY);
        c.t_hi = dev_tap_base(jb, dY) + 3;
        c.new_lo = k == 0 ? c.t_lo : max(c.t_lo, dev_tap_base(ja - 1, dY) + 4);
    } else { // COPY tiles address rows relative to the tile's first row (keeps everything in int)
        c.t_lo = k * RC;
        c.t_hi = (int)(jb - T.j0);
        c.new_lo = c.t_lo;
    }
    c.n_new = max(0, c.t_hi - c.new_lo + 1);
    return c;
}

// ring slots s..s+4 of one row, each an (left-half pixel, right-half pixel) pair
struct RowRegs {
    f2 S[5];
};
__device__ __forceinline__ void load_row_regs(RowRegs &R, const f2 *p)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(R.S[0]), "=l"(R.S[1]) : "r"(smem_u32(p)));
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(R.S[2]), "=l"(R.S[3]) : "r"(smem_u32(p + 2)));
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(R.S[4]) : "r"(smem_u32(p + 4)));
}
// per-row dot product in OpenCV's interior order ((s0*w0 + s1*w1) + s2*w2) + s3*w3 for output slot o (0 or 1)
template <int O>
__device__ __forceinline__ f2 dot_row(const RowRegs &R, const f2 (&W)[4], f2 nz)
{
    return add2(add2(add2(mul2(R.S[O], W[0], nz), mul2(R.S[O + 1], W[1], nz)), mul2(R.S[O + 2], W[2], nz)),
                mul2(R.S[O + 3], W[3], nz));
}

__device__ __forceinline__ uint32_t cast_u16(float s)
{
    return (uint32_t)max(0, min(65535, __float2int_rn(s))); // cvRound + saturate_cast<ushort>
}

// ring addressing: source-window column sc (0..SRC_W) -> float index inside a ring row
__device__ __forceinline__ int ring_idx(int sc) { return sc < SLOTS ? 2 * sc : 2 * (sc - HALF) + 1; }

// generic single-pixel bicubic from the ring (any alignment, border or interior order)
struct RingView {
    const float *ring;
    int t_base; // buffer row held in ring row 0 (mod RING)
    int ix0;    // source column of window column 0
    int t_lo, t_hi;
};
__device__ __forceinline__ uint32_t resample_px_general(const RingView &V, const float *s_tab, int sx, int sy, int w,
                                                        int hbuf)
{
    const int ix = dev_sat_short(sx >> 5) - 1, fx = sx & 31;
    const int iy = dev_sat_short(sy >> 5) - 1, fy = sy & 31;
    float v[4][4], wg[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int t = iy + r;
        const bool rin = t >= V.t_lo && t <= V.t_hi;
        const int slot = rin ? (t - V.t_base) % RING : 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int sc = ix + c - V.ix0;
            const bool in = rin && sc >= 0 && sc < SRC_W;
            v[r][c] = in ? V.ring[slot * (2 * SLOTS) + ring_idx(sc)] : 0.f;
            wg[r][c] = __fmul_rn(s_tab[4 * fy + r], s_tab[4 * fx + c]);
        }
    }
    const bool interior = (unsigned)ix < (unsigned)max(w - 3, 0) && (unsigned)iy < (unsigned)max(hbuf - 3, 0);
    float s;
    if (interior) {
        s = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[r][0], wg[r][0]), __fmul_rn(v[r][1], wg[r][1])),
                                          __fmul_rn(v[r][2], wg[r][2])),
                                __fmul_rn(v[r][3], wg[r][3]));
            s = r == 0 ? d : __fadd_rn(s, d);
        }
    } else { // border: flat accumulation from 0, out-of-image taps contribute 0 (they are 0 in the ring)
        s = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) s = __fadd_rn(s, __fmul_rn(v[r][c], wg[r][c]));
    }
    return cast_u16(s);
}

// ---------------------------------------------------------------------------------------------
// convert helpers: staged halfwords -> zero-extended (byte-swapped) samples, 8-wide fp64 RRC
// ---------------------------------------------------------------------------------------------
// N samples starting at halfword (ODD ? 1 : 0) of wd[]; one PRMT does select + swap + zero extension
template <bool SWAP, bool ODD, int N, int NW>
__device__ __forceinline__ void extract(const uint32_t (&wd)[NW], uint32_t *s)
{
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int h = (ODD ? 1 : 0) + j;
        const uint32_t sel = (h & 1) ? (SWAP ? 0x4423u : 0x4432u) : (SWAP ? 0x4401u : 0x4410u);
        s[j] = __byte_perm(wd[h >> 1], 0u, sel);
    }
}

__device__ __forceinline__ void rrc8(uint32_t (&s)[8], const double (&k)[8], const double (&b)[8])
{
    double v[8];
    uint32_t hmax = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // u16 -> f64: one I2F.F64.U32 here (conversion pipe, 16 lanes/clk/SM) costs fewer issue slots than
        // the magic-number form (MOV hi + DADD) and this kernel is issue-bound, not conversion-pipe-bound
        const double sd = __uint2double_rn(s[j]);
        v[j] = __dadd_rn(__dmul_rn(k[j], sd), b[j]);
        hmax = max(hmax, (uint32_t)__double2hiint(v[j]));
    }
    // every value of the warp in [0, 2^31): exact truncation with one DADD.RZ each (warp-uniform branch)
    if (!__any_sync(__activemask(), hmax >= 0x41E00000u)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = (uint32_t)__double2loint(__dadd_rz(v[j], 4503599627370496.0));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = (v[j] > -2147483649.0 && v[j] < 2147483648.0) ? __double2int_rz(v[j]) : (int)0x80000000;
            s[j] = (uint32_t)t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// convert step: one thread-item = 8 samples of one staged row -> swap -> RRC -> ring / output.
// REMAP: window columns 4q..4q+3 (lane 0) and 4q+HALF.. (lane 1) -> 4 ring slots (two 128-bit stores)
// COPY : tile columns 8q..8q+7 -> 128-bit store straight into the output raster
// SWAP / ODD (staging halfword parity) / REMAP are compile-time so the row loop has no dispatch.
// ---------------------------------------------------------------------------------------------
struct ConvertCtx {
    const uint16_t *buf;   // staging buffer of this chunk
    const uint8_t *rz;     // per staged row: 1 = zero border row
    float *ring;
    uint16_t *out_rows;    // COPY: output pointer of (tile row 0 of this chunk's new_lo, column 8q)
    int64_t out_pitch;
    int word0;             // first staging word of my samples
    int slot0;             // REMAP: ring slot of staged row 0
    int n_new, ph, q, n_cols;
    bool need_rrc, copy_vec;
};

template <bool REMAP, bool SWAP, bool ODD>
__device__ __forceinline__ void convert_rows(const ConvertCtx &X, const double (&kk)[8], const double (&bb)[8])
{
    int slot = X.slot0 + X.ph;
    if (slot >= RING) slot -= RING;
    for (int r = X.ph; r < X.n_new; r += CV_PH) {
        uint32_t s[8];
        if (X.rz[r] == 0) {
            const uint32_t *row32 = reinterpret_cast<const uint32_t *>(X.buf + (size_t)r * SWC) + X.word0;
            if (REMAP) {
                const uint32_t *wb = row32 + HALF / 2;
                uint32_t a[3] = {row32[0], row32[1], ODD ? row32[2] : 0u};
                uint32_t b[3] = {wb[0], wb[1], ODD ? wb[2] : 0u};
                extract<SWAP, ODD, 4, 3>(a, s);
                extract<SWAP, ODD, 4, 3>(b, s + 4);
            } else {
                uint32_t wd[5] = {row32[0], row32[1], row32[2], row32[3], ODD ? row32[4] : 0u};
                extract<SWAP, ODD, 8, 5>(wd, s);
            }
            if (X.need_rrc) rrc8(s, kk, bb);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = 0;
        }
        if (REMAP) {
            float4 o0, o1; // (slot 4q: L,R) (slot 4q+1: L,R) | (slot 4q+2) (slot 4q+3)
            // low 16 bits -> exact float: PRMT builds 0x4B00hhll (= 2^23 + v), one FADD removes the 2^23
            auto f = [](uint32_t v) { return __fadd_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7610)), -8388608.0f); };
            o0.x = f(s[0]); o0.y = f(s[4]); o0.z = f(s[1]); o0.w = f(s[5]);
            o1.x = f(s[2]); o1.y = f(s[6]); o1.z = f(s[3]); o1.w = f(s[7]);
            float4 *dst = reinterpret_cast<float4 *>(X.ring + (size_t)slot * (2 * SLOTS) + 8 * X.q);
            dst[0] = o0;
            dst[1] = o1;
            if (slot < RING_MIRROR) { // mirrored copy: any run of RING_MIRROR+1 ring rows is contiguous
                float4 *dm = dst + (size_t)RING * (2 * SLOTS) / 4;
                dm[0] = o0;
                dm[1] = o1;
            }
            slot += CV_PH;
            if (slot >= RING) slot -= RING;
        } else {
            uint16_t *orow = X.out_rows + (int64_t)r * X.out_pitch;
            if (X.copy_vec && 8 * X.q + 8 <= X.n_cols) {
                uint4 o;
                o.x = __byte_perm(s[0], s[1], 0x5410);
                o.y = __byte_perm(s[2], s[3], 0x5410);
                o.z = __byte_perm(s[4], s[5], 0x5410);
                o.w = __byte_perm(s[6], s[7], 0x5410);
                stg_na_v4(orow, o);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (8 * X.q + j < X.n_cols) orow[j] = (uint16_t)s[j];
            }
        }
    }
}

template <bool REMAP>
__device__ __forceinline__ void convert_dispatch(const ConvertCtx &X, bool swap, bool odd, const double (&kk)[8],
                                                 const double (&bb)[8])
{
    if (swap) {
        if (odd) convert_rows<REMAP, true, true>(X, kk, bb); else convert_rows<REMAP, true, false>(X, kk, bb);
    } else {
        if (odd) convert_rows<REMAP, false, true>(X, kk, bb); else convert_rows<REMAP, false, false>(X, kk, bb);
    }
}

__global__ void __launch_bounds__(NT, CTAS_PER_SM) pan_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    float *ring = reinterpret_cast<float *>(smem);                                            // RING_ROWS x SLOTS float2
    uint16_t *stg = reinterpret_cast<uint16_t *>(smem + (size_t)RING_ROWS * SLOTS * 8);       // 2 x STG x SWC u16
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ __align__(8) float s_tab[132]; // 32x4 weights + the (-0.0,-0.0) pair at [128..129]
    __shared__ uint8_t s_rowzero[2][STG];

    const int tid = threadIdx.x;
    const Tile T = P.tiles[blockIdx.x];
    const CcdDev &C = P.ccd[T.ccd];
    const int w = P.w;
    const bool remap = T.kind == KIND_REMAP;
    const bool bulk = P.bulk_ok != 0;
    const bool swap = bulk && C.fmt == OIP_FMT_BE16; // the generic loader already delivers native order
    const bool do_rrc = C.kb != nullptr;
    const double dX = C.dX, dY = C.dY;

    for (int i = tid; i < 130; i += NT) s_tab[i] = P.tab[i];
    if (tid == 0) {
        mbar_init(&bars[0], NWARPS); // one arrive.expect_tx per warp (each warp issues its share of the rows)
        mbar_init(&bars[1], NWARPS);
        fence_mbar_init();
    }

    // ---- column geometry: window column 0 <-> source column ix0; staging column 0 <-> c_lo
    const int n_cols = T.x_end - T.x_begin;
    const int sx0 = remap ? dev_map_fixed(T.x_begin, dX) : 0;
    const int ix0 = remap ? dev_sat_short(sx0 >> 5) - 1 : T.x_begin;
    const int c_lo = (ix0 >= 0 ? ix0 : ix0 - 7) / 8 * 8; // floor to a multiple of 8 (16-byte aligned bulk copies)
    const int delta = ix0 - c_lo;
    const int n_chunks = (T.n_rows + RC - 1) / RC;

    // ---- regular tile: the fixed-point maps advance by exactly one source pixel per output pixel /
    //      row over the whole tile (always, except where float(i + d) rounds across a 1/32 boundary).
    //      Then every per-chunk quantity is plain integer arithmetic.
    const int sy_t0 = remap ? dev_map_fixed(T.j0, dY) : 0;
    bool mine_regular = true;
    if (remap) {
        if (tid < n_cols) mine_regular = dev_map_fixed(T.x_begin + tid, dX) == sx0 + 32 * tid;
        for (int i = tid; i < T.n_rows; i += NT) mine_regular = mine_regular && (dev_map_fixed(T.j0 + i, dY) == sy_t0 + 32 * i);
    }
    const bool regular = __syncthreads_and(mine_regular) != 0; // also publishes s_tab / barriers
    const int iy_t0 = (sy_t0 >> 5) - 1;                         // first tap row of tile row 0 (regular tiles)

    auto get_chunk = [&](int k) {
        if (remap && regular) {
            ChunkRows c;
            const int nr = min(RC, T.n_rows - k * RC);
            c.t_lo = iy_t0 + k * RC;
            c.t_hi = c.t_lo + nr + 2;
            c.new_lo = k == 0 ? c.t_lo : c.t_lo + 3;
            c.n_new = c.t_hi - c.new_lo + 1;
            return c;
        }
        return chunk_rows(T, dY, k);
    };
    const int t_base = get_chunk(0).t_lo;

    // ---- convert-step mapping and RRC coefficients (registers)
    const int cv_q = tid & (CV_Q - 1), cv_ph = tid / CV_Q;
    double kk[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ix0 + (remap ? (4 * cv_q + (j & 3) + (j >> 2) * HALF) : (8 * cv_q + j));
        const bool ok = c >= 0 && c < w;
        kk[j] = (ok && do_rrc) ? C.kb[2 * c] : (ok ? 1.0 : 0.0); // k = b = 0 zeroes columns outside the CCD
        bb[j] = (ok && do_rrc) ? C.kb[2 * c + 1] : 0.0;
    }
    const bool need_rrc = do_rrc || ix0 < 0 || ix0 + SRC_W > w; // uniform: RRC, or some window column is outside the CCD
    const bool cv_active = remap ? true : (8 * cv_q < n_cols);
    uint16_t *const out_tile = P.out + (T.g0 - P.out_row0) * P.out_pitch + T.out_x;
    const bool copy_vec = ((T.out_x & 7) == 0) && ((P.out_pitch & 7) == 0) && ((((uintptr_t)P.out) & 15) == 0);

    // ------------------------------------------------------------------ loader
    auto issue_chunk = [&](int k) {
        const ChunkRows cr = get_chunk(k);
        uint16_t *buf = stg + (size_t)(k & 1) * STG * SWC;
        uint8_t *rz = s_rowzero[k & 1];
        uint64_t *bar = &bars[k & 1];
        const int64_t row_bias = remap ? 0 : T.j0; // COPY rows are tile-relative
        if (bulk) {
            // every warp issues rows wid, wid+NWARPS, ... (lane i takes the i-th of them): the issue
            // loop is serialised per lane by the uniform datapath, so spreading it keeps one warp
            // from becoming the straggler at the next barrier
            const int lane = tid & 31, wid = tid >> 5;
            const int ca = max(c_lo, 0), cb = min(c_lo + SWC, w);
            const uint32_t nb = cb > ca ? (uint32_t)(cb - ca) * 2u : 0u;
            const int r = wid + NWARPS * lane;
            // common case: the chunk's rows are consecutive lines of one segment
            const int64_t ga = local_to_global(T, row_bias + cr.new_lo);
            const int64_t gb = local_to_global(T, row_bias + cr.new_lo + cr.n_new - 1);
            const uint8_t *run = nullptr;
            int64_t run_pitch = 0;
            if (cr.n_new > 0 && ga >= 0 && gb - ga == cr.n_new - 1 && nb) {
#pragma unroll
                for (int s = 0; s < OIP_MAX_SEG; ++s)
                    if (s < C.n_seg && ga >= C.seg[s].row0 && gb < C.seg[s].row0 + C.seg[s].n_rows) {
                        run = C.seg[s].base + (ga - C.seg[s].row0) * C.seg[s].pitch;
                        run_pitch = C.seg[s].pitch;
                    }
            }
            const uint8_t *q = nullptr;
            if (r < cr.n_new) {
                if (run) {
                    q = run + r * run_pitch;
                } else {
                    const int64_t g = local_to_global(T, row_bias + cr.new_lo + r);
                    q = (g >= 0 && nb) ? row_ptr(C, g) : nullptr;
                    if (g >= 0 && nb && !q) atomicExch(P.err, 1); // host failed to supply a needed row
                }
                rz[r] = q == nullptr;
            }
            uint32_t total = q ? nb : 0u;
#pragma unroll
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) mbar_arrive_expect_tx(bar, total);
            __syncwarp();
            if (q) bulk_g2s(buf + (size_t)r * SWC + (ca - c_lo), q + 2 * (int64_t)ca, nb, bar);
        } else {
            // generic loader: any alignment / packed / tile layouts; native byte order in staging
            for (int r = 0; r < cr.n_new; ++r) {
                const int64_t g = local_to_global(T, row_bias + cr.new_lo + r);
                const uint8_t *q = nullptr;
                if (g >= 0) q = C.fmt == OIP_FMT_BE16_TILES ? C.seg[0].base : row_ptr(C, g);
                if (g >= 0 && !q && tid == 0) atomicExch(P.err, 1);
                if (tid == 0) rz[r] = q == nullptr;
                if (q)
                    for (int i = tid; i < SWC; i += NT) {
                        const int c = c_lo + i;
                        buf[(size_t)r * SWC + i] = (c >= 0 && c < w) ? (uint16_t)load_sample(C, q, g, c) : (uint16_t)0;
                    }
            }
        }
    };

    issue_chunk(0);
    if (n_chunks > 1) issue_chunk(1);
    __syncthreads();

    // ---- resample-step mapping: thread = (slot pair cg, row group rg) -> slots 2cg,2cg+1 x 8 rows
    //      = tile columns {2cg, 2cg+1} (lane 0) and {HALF+2cg, HALF+2cg+1} (lane 1)
    const int cg = tid & (RS_CG - 1), rg = tid / RS_CG;
    const bool rs_active = cg < HALF / 2;
    const int xl = 2 * cg, xr = HALF + 2 * cg; // my left / right column pairs
    // pixels that exist and whose 4x4 footprint is inside the CCD may use the packed interior path
    auto col_interior = [&](int xi) { return (ix0 + xi) >= 0 && (ix0 + xi) < w - 3; };
    const bool l_exists = xl < n_cols, r_exists = xr < n_cols; // pairs exist or not as a whole when n_cols is even
    const bool fast_ok = rs_active && regular && ((n_cols & 1) == 0) && l_exists && col_interior(xl) &&
                         col_interior(xl + 1) && (!r_exists || (col_interior(xr) && col_interior(xr + 1)));
    const int fx = sx0 & 31, fy = sy_t0 & 31;
    const bool out_vec2 = ((((uintptr_t)P.out) & 3) == 0) && ((P.out_pitch & 1) == 0) && ((T.out_x & 1) == 0);

    for (int k = 0; k < n_chunks; ++k) {
        const ChunkRows cr = get_chunk(k);
        const int nr = min(RC, T.n_rows - k * RC);
        if (bulk) mbar_wait(&bars[k & 1], (uint32_t)((k >> 1) & 1));

        // ---------------------------------------------------------- convert: swap + RRC
        if (cv_active) {
            ConvertCtx X;
            X.buf = stg + (size_t)(k & 1) * STG * SWC;
            X.rz = s_rowzero[k & 1];
            X.ring = ring;
            X.out_rows = out_tile + (int64_t)cr.new_lo * P.out_pitch + 8 * cv_q;
            X.out_pitch = P.out_pitch;
            X.word0 = (delta + (remap ? 4 : 8) * cv_q) >> 1;
            X.slot0 = remap ? (cr.new_lo - t_base) % RING : 0;
            X.n_new = cr.n_new; X.ph = cv_ph; X.q = cv_q; X.n_cols = n_cols;
            X.need_rrc = need_rrc; X.copy_vec = copy_vec;
            if (remap) convert_dispatch<true>(X, swap, (delta & 1) != 0, kk, bb);
            else convert_dispatch<false>(X, swap, (delta & 1) != 0, kk, bb);
        }
        __syncthreads();
        if (k + 2 < n_chunks) issue_chunk(k + 2);

        // ---------------------------------------------------------- resample
        if (remap && rs_active) {
            uint16_t *o = out_tile + ((int64_t)k * RC + 8 * rg) * P.out_pitch;
            const int iy0 = iy_t0 + k * RC; // regular tiles: first tap row of the chunk
            if (fast_ok && nr == RC && iy0 >= 0 && iy0 + RC - 1 < T.hbuf - 3) {
                const f2 nz = *reinterpret_cast<const f2 *>(&s_tab[128]);
                f2 W[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float wv = __fmul_rn(s_tab[4 * fy + r], s_tab[4 * fx + c]);
                        W[r][c] = pk(wv, wv);
                    }
                // 11 consecutive ring rows from slot0: contiguous thanks to the mirrored rows
                const int slot0 = (iy0 + 8 * rg - t_base) % RING;
                const f2 *rowp = reinterpret_cast<const f2 *>(ring) + (size_t)slot0 * SLOTS + 2 * cg;
                RowRegs R0, R1, R2, R3;
                auto emit = [&](const RowRegs &a, const RowRegs &b, const RowRegs &c, const RowRegs &d) {
                    f2 p0 = dot_row<0>(a, W[0], nz), p1 = dot_row<1>(a, W[0], nz);
                    p0 = add2(p0, dot_row<0>(b, W[1], nz)); p1 = add2(p1, dot_row<1>(b, W[1], nz));
                    p0 = add2(p0, dot_row<0>(c, W[2], nz)); p1 = add2(p1, dot_row<1>(c, W[2], nz));
                    p0 = add2(p0, dot_row<0>(d, W[3], nz)); p1 = add2(p1, dot_row<1>(d, W[3], nz));
                    const uint32_t l0 = cast_u16(lo_of(p0)), l1 = cast_u16(lo_of(p1));
                    const uint32_t r0 = cast_u16(hi_of(p0)), r1 = cast_u16(hi_of(p1));
                    if (out_vec2) {
                        *reinterpret_cast<uint32_t *>(o + xl) = l0 | (l1 << 16);
                        if (r_exists) *reinterpret_cast<uint32_t *>(o + xr) = r0 | (r1 << 16);
                    } else {
                        o[xl] = (uint16_t)l0; o[xl + 1] = (uint16_t)l1;
                        if (r_exists) { o[xr] = (uint16_t)r0; o[xr + 1] = (uint16_t)r1; }
                    }
                    o += P.out_pitch;
                };
                load_row_regs(R0, rowp);
                load_row_regs(R1, rowp + SLOTS);
                load_row_regs(R2, rowp + 2 * SLOTS);
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    load_row_regs(R3, rowp + 3 * SLOTS); emit(R0, R1, R2, R3);
                    load_row_regs(R0, rowp + 4 * SLOTS); emit(R1, R2, R3, R0);
                    load_row_regs(R1, rowp + 5 * SLOTS); emit(R2, R3, R0, R1);
                    load_row_regs(R2, rowp + 6 * SLOTS); emit(R3, R0, R1, R2);
                    rowp += 4 * SLOTS;
                }
            } else {
                // exact generic path: borders, section edges, partial chunks, odd widths, map anomalies
                RingView V{ring, t_base, ix0, cr.t_lo, cr.t_hi};
                for (int i = 0; i < 8; ++i) {
                    const int row = 8 * rg + i;
                    if (row < nr) {
                        const int64_t j = T.j0 + (int64_t)k * RC + row;
                        const int sy = regular ? sy_t0 + 32 * (k * RC + row) : dev_map_fixed(j, dY);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int xi = (c < 2 ? xl : xr) + (c & 1);
                            if (xi < n_cols) {
                                const int sx = regular ? sx0 + 32 * xi : dev_map_fixed(T.x_begin + xi, dX);
                                o[xi] = (uint16_t)resample_px_general(V, s_tab, sx, sy, w, T.hbuf);
                            }
                        }
                    }
                    o += P.out_pitch;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ host side
struct PlanKey {
    int n_ccd, w, fold_half, section_rows, row_guard, shifted[8];
    int64_t total_rows, row0, n_rows, out_pitch;
    double dX[8], dY[8];
    // what the fast-path split depends on: eligibility per CCD, row-segment layout, tunables
    int fast[8], n_seg[8], fast_rows, fmt[8], tile_cols[8], tile_lines[8];
    int64_t seg_row0[8][OIP_MAX_SEG], seg_rows[8][OIP_MAX_SEG];
};

// generic tiles (pan_kernel) over the rectangle [x_lo,x_hi) x [g_lo,g_hi) of CCD i inside shift segment s
static void emit_generic(std::vector<Tile> &tiles, int i, bool shifted, const ShiftSegment &s, int x_lo, int x_hi,
                         int out_x_lo, int64_t g_lo, int64_t g_hi)
{
    if (x_hi <= x_lo || g_hi <= g_lo) return;
    // slivers (leftover columns next to the fast spans) get somewhat shorter tiles; the launch runs on a side
    // stream next to the fast kernel, so a few long CTAs are better than thousands of tiny ones
    const int th = x_hi - x_lo < 64 ? 256 : TH;
    for (int64_t g = g_lo; g < g_hi; g += th) {
        const int nr = (int)std::min<int64_t>(th, g_hi - g);
        // column strips of equal width (no skinny last strip); even for the 32-bit paired stores of REMAP
        // tiles, multiple of 8 for the 128-bit stores of COPY tiles
        const int n_strips = (x_hi - x_lo + TW - 1) / TW;
        const int align = shifted ? 2 : 8;
        const int sw = std::min(TW, ((x_hi - x_lo + n_strips - 1) / n_strips + align - 1) / align * align);
        for (int x = x_lo; x < x_hi; x += sw) {
            Tile t{};
            t.ccd = i;
            t.kind = shifted ? KIND_REMAP : KIND_COPY;
            t.x_begin = x;
            t.x_end = std::min(x_hi, x + sw);
            t.out_x = out_x_lo + (x - x_lo);
            t.n_rows = nr;
            t.g0 = g;
            t.j0 = shifted ? s.j0 + (g - s.g0) : g;
            t.sec_off = s.sec_off;
            t.stale_off = s.stale_off;
            t.rows_s = s.rows_s;
            t.hbuf = s.hbuf;
            tiles.push_back(t);
        }
    }
}

// index of the row segment of CCD c that holds all of the global source rows [a, b], -1 if none does
static int seg_holding(const oip_ccd_src &c, int64_t a, int64_t b)
{
    for (int k = 0; k < c.n_seg; ++k)
        if (a >= c.seg[k].row0 && b < c.seg[k].row0 + c.seg[k].n_rows) return k;
    return -1;
}

struct ColSpan { int xa, xb, sx_a; }; // columns [xa,xb) on the fast path; sx_a = fixed-point map of column xa

// warp-tiles (pan_fast_kernel) for output rows [ga,gb) x the column span; the first source row of output row
// ga is row `src_row_a` of row segment `seg`; fx/fy = sub-pixel phase
static void emit_fast(std::vector<panfast::FastTile> &out, int kind, int i, int seg, const ColSpan &cs, int lo, int out_x_ccd,
                      int64_t ga, int64_t gb, int64_t src_row_a, int fy, const oip_pan_desc *d, int th)
{
    const int max_w = kind == panfast::FT_REMAP ? 2 * panfast::HALF_MAX : panfast::COPY_MAX;
    const int width = cs.xb - cs.xa; // REMAP: multiple of 8; COPY: any (the last strip takes the odd columns)
    const int n_strips = (width + max_w - 1) / max_w;
    const int base = width / n_strips / 8 * 8;       // < max_w unless width == n_strips * max_w: the +8 / tail below fit
    int extra = (width - base * n_strips) / 8;       // this many strips are 8 columns wider
    const int tail = (width - base * n_strips) % 8;  // COPY only: odd columns, given to the last strip
    const int64_t len = gb - ga;
    const int n_t = (int)((len + th - 1) / th);
    const int64_t h = (len + n_t - 1) / n_t;
    for (int64_t g = ga; g < gb; g += h) {
        int x = cs.xa, ex = extra;
        for (int sidx = 0; sidx < n_strips; ++sidx) {
            const int sw = base + (ex > 0 ? 8 : 0) + (sidx == n_strips - 1 ? tail : 0);
            if (ex > 0) --ex;
            panfast::FastTile t{};
            t.kind = kind;
            t.ccd = i;
            t.tmap = i * OIP_MAX_SEG + seg;
            t.x_begin = x;
            t.n_rows = (int)std::min<int64_t>(h, gb - g);
            t.src_y0 = (int)(src_row_a + (g - ga));
            t.out_off = (g - d->row0) * d->out_pitch_px + out_x_ccd + (x - lo);
            if (kind == panfast::FT_REMAP) {
                const int sx = cs.sx_a + 32 * (x - cs.xa);
                t.half = sw / 2;
                t.src_x0 = sat_short(sx >> 5) - 1;
                t.fx = sx & 31;
                t.fy = fy;
            } else {
                t.half = sw;
                t.src_x0 = x;
            }
            out.push_back(t);
            x += sw;
        }
    }
}

// column gap [xa,xb) of a shifted CCD over the fast row run [ga,gb): EDGE warp-tiles (one column per lane, both
// accumulation orders, any alignment) where the fixed-point map is regular across the gap, generic tiles otherwise
static void emit_gap(std::vector<Tile> &tiles, std::vector<panfast::FastTile> &fl, int i, int seg, const ShiftSegment &s, int xa, int xb,
                     int lo, int out_x_ccd, int64_t ga, int64_t gb, int64_t src_row_a, int fy, const oip_pan_desc *d, int th)
{
    if (xb <= xa) return;
    const double dX = d->ccd[i].dX;
    int x = xa;
    while (x < xb) {
        const int sx0 = map_fixed(x, dX);
        int xe = x + 1;
        while (xe < xb && xe - x < panfast::EDGE_MAX && map_fixed(xe, dX) == sx0 + 32 * (xe - x)) ++xe;
        const int ix0 = (sx0 >> 5) - 1;
        if (ix0 < -30000 || ix0 > 30000) { // saturating coordinates: not a case for the fast path
            emit_generic(tiles, i, true, s, x, xe, out_x_ccd + (x - lo), ga, gb);
            x = xe;
            continue;
        }
        const int64_t len = gb - ga;
        const int n_t = (int)((len + th - 1) / th);
        const int64_t h = (len + n_t - 1) / n_t;
        for (int64_t g = ga; g < gb; g += h) {
            panfast::FastTile t{};
            t.kind = panfast::FT_EDGE;
            t.ccd = i;
            t.tmap = i * OIP_MAX_SEG + seg;
            t.x_begin = x;
            t.half = xe - x;
            t.src_x0 = ix0;
            t.src_y0 = (int)(src_row_a + (g - ga));
            t.n_rows = (int)std::min<int64_t>(h, gb - g);
            t.fx = sx0 & 31;
            t.fy = fy;
            t.out_off = (g - d->row0) * d->out_pitch_px + out_x_ccd + (x - lo);
            fl.push_back(t);
        }
        x = xe;
    }
}

static int build_plan(const oip_pan_desc *d, const bool *fast_ccd, int fast_rows, std::vector<Tile> &tiles,
                      std::vector<panfast::FastTile> &ftiles, int64_t (&cls_ctas)[3])
{
    const int n = d->n_ccd, w = d->w, f = d->fold_half;
    const int64_t r_lo = d->row0, r_hi = d->row0 + d->n_rows;
    std::vector<panfast::FastTile> fl;
    // keep the plan small on very long strips: taller warp-tiles
    int th = std::max(16, fast_rows);
    if (fast_rows == 128 && d->n_rows >= 65536) th = 256; // long strips: the per-tile prologue amortises better (measured, tools/sweep_sizes.py)
    while ((double)d->n_rows / th * ((double)n * w / 248.0) > 400000.0) th *= 2;
    while (th > 16 && (int64_t)(th + 8) * d->out_pitch_px * 2 >= (1ll << 32)) th /= 2; // REMAP tiles step a 32-bit row offset (pan_fast.cu)
    int out_x = 0;
    for (int i = 0; i < n; ++i) {
        const oip_ccd_src &C = d->ccd[i];
        const int lo = i == 0 ? 0 : f, hi = i == n - 1 ? w : w - f;
        const bool shifted = C.shifted != 0, fast = fast_ccd[i];
        if (!shifted) {
            const ShiftSegment s{0, d->total_rows, 0, 0, 0, -1, 0};
            // COPY: TMA box origins need source columns that are multiples of 8 (16 bytes), the 128-bit stores need
            // the same of the output column
            const int xs = (lo + 7) & ~7;
            const int usable = fast && xs < hi && (out_x + xs - lo) % 8 == 0 ? hi - xs : 0;
            int64_t cur = r_lo;
            if (usable >= 8) {
                const ColSpan cs{xs, xs + usable, 0};
                while (cur < r_hi) {
                    const int k = seg_holding(C, cur, cur);
                    if (k < 0) break;
                    const int64_t end = std::min<int64_t>(r_hi, C.seg[k].row0 + C.seg[k].n_rows);
                    emit_fast(fl, panfast::FT_COPY, i, k, cs, lo, out_x, cur, end, cur - C.seg[k].row0, 0, d, th);
                    emit_generic(tiles, i, false, s, lo, xs, out_x, cur, end);
                    emit_generic(tiles, i, false, s, xs + usable, hi, out_x + (xs + usable - lo), cur, end);
                    cur = end;
                }
            }
            emit_generic(tiles, i, false, s, lo, hi, out_x, cur, r_hi); // rows no single segment holds (or no fast path)
            out_x += hi - lo;
            continue;
        }
        std::vector<ShiftSegment> segs;
        if (!plan_shift_segments(d->total_rows, d->section_rows, d->row_guard, C.dY, segs))
            return fail(OIP_E_INVALID, "invalid section geometry (section_rows=%d row_guard=%d dY=%g)", d->section_rows,
                        d->row_guard, C.dY);
        // fast column spans: interior footprints, regular map, output column a multiple of 4 (64-bit stores)
        std::vector<ColSpan> spans;
        if (fast) {
            int x = lo;
            while (x < hi) {
                const int sx = map_fixed(x, C.dX), ix = sat_short(sx >> 5) - 1;
                if (!(ix >= 0 && ix < w - 3)) { ++x; continue; }
                int xe = x + 1, sxe = sx;
                while (xe < hi) {
                    const int s2 = map_fixed(xe, C.dX), i2 = sat_short(s2 >> 5) - 1;
                    if (s2 != sxe + 32 || !(i2 >= 0 && i2 < w - 3)) break;
                    sxe = s2;
                    ++xe;
                }
                const int xs = x + ((4 - ((out_x + x - lo) % 4)) % 4);
                const int usable = xs < xe ? ((xe - xs) & ~7) : 0;
                if (usable >= 16) spans.push_back({xs, xs + usable, sx + 32 * (xs - x)});
                x = xe;
            }
        }
        for (const ShiftSegment &s : segs) {
            const int64_t ga = std::max<int64_t>(s.g0, r_lo), gb = std::min<int64_t>(s.g1, r_hi);
            if (gb <= ga) continue;
            if (spans.empty()) {
                emit_generic(tiles, i, true, s, lo, hi, out_x, ga, gb);
                continue;
            }
            // fast row runs: 4 tap rows inside the section's fresh rows, regular map, one row segment
            int64_t g = ga, gen_from = ga;
            while (g < gb) {
                auto probe = [&](int64_t gg, int &sy, int &k) {
                    const int64_t j = s.j0 + (gg - s.g0);
                    sy = map_fixed(j, C.dY);
                    const int t = sat_short(sy >> 5) - 1;
                    if (!(t >= 0 && t + 3 < s.rows_s && t + 3 < s.hbuf)) return false;
                    k = seg_holding(C, s.sec_off + t, s.sec_off + t + 3);
                    return k >= 0;
                };
                int sy0, k0;
                if (!probe(g, sy0, k0)) { ++g; continue; }
                int64_t ge = g + 1;
                int syp = sy0;
                while (ge < gb) {
                    int sy, k;
                    if (!probe(ge, sy, k) || sy != syp + 32 || k != k0) break;
                    syp = sy;
                    ++ge;
                }
                if (ge - g >= 8) {
                    emit_generic(tiles, i, true, s, lo, hi, out_x, gen_from, g); // rows before the run: full width
                    const int64_t src_row = s.sec_off + (sat_short(sy0 >> 5) - 1) - C.seg[k0].row0;
                    int xg = lo; // columns between the fast spans stay generic
                    for (const ColSpan &cs : spans) {
                        emit_gap(tiles, fl, i, k0, s, xg, cs.xa, lo, out_x, g, ge, src_row, sy0 & 31, d, th);
                        emit_fast(fl, panfast::FT_REMAP, i, k0, cs, lo, out_x, g, ge, src_row, sy0 & 31, d, th);
                        xg = cs.xb;
                    }
                    emit_gap(tiles, fl, i, k0, s, xg, hi, lo, out_x, g, ge, src_row, sy0 & 31, d, th);
                    gen_from = ge;
                }
                g = ge;
            }
            emit_generic(tiles, i, true, s, lo, hi, out_x, gen_from, gb);
        }
        out_x += hi - lo;
    }
    // CTA = WARPS consecutive warp-tiles (its warps are independent workers), row bands in raster order:
    // neighbouring strips share halo columns through L2, COPY and REMAP tiles interleave on every SM
    // (putting all long REMAP tiles first and the short COPY tiles at the end of the launch was measured: no difference)
    std::stable_sort(fl.begin(), fl.end(), [&](const panfast::FastTile &a, const panfast::FastTile &b) {
        const int ca = panfast::fast_class(d->ccd[a.ccd].fmt), cb = panfast::fast_class(d->ccd[b.ccd].fmt);
        if (ca != cb) return ca < cb; // one launch per source-format class (a strip is almost always of one class)
        const int64_t ra = a.out_off / d->out_pitch_px / th, rb = b.out_off / d->out_pitch_px / th;
        if (ra != rb) return ra < rb;
        if (a.kind != b.kind) return a.kind > b.kind; // REMAP tiles (long) first inside a band
        if (a.ccd != b.ccd) return a.ccd < b.ccd;
        return a.x_begin < b.x_begin;
    });
    for (panfast::FastTile &t : fl) { // tiled sources: the sub-image column of the stage box origin instead of a tensor map
        const oip_ccd_src &C = d->ccd[t.ccd];
        if (C.fmt != OIP_FMT_BE16_TILES) continue;
        const int x0 = t.kind == panfast::FT_COPY ? t.x_begin : (t.src_x0 & ~7);
        t.tmap = x0 >= 0 ? x0 / C.tile_cols : -((-x0 + C.tile_cols - 1) / C.tile_cols);
    }
    panfast::FastTile none{};
    none.kind = panfast::FT_NONE;
    ftiles.clear();
    for (int c = 0; c < 3; ++c) { // every class padded to whole CTAs
        const size_t before = ftiles.size();
        for (const panfast::FastTile &t : fl)
            if (panfast::fast_class(d->ccd[t.ccd].fmt) == c) ftiles.push_back(t);
        while (ftiles.size() % panfast::WARPS) ftiles.push_back(none);
        cls_ctas[c] = (int64_t)((ftiles.size() - before) / panfast::WARPS);
    }
    return OIP_OK;
}

static float g_tab_host[128];
static bool g_tab_ready = false;
static void cubic_tab_host(float *tab)
{
    // OpenCV interpolateCubic, A = -0.75, evaluated in float without contraction (host code of
    // this file is compiled by the host compiler for baseline x86-64: no FMA instructions)
    const volatile float A = -0.75f;
    for (int i = 0; i < 32; ++i) {
        volatile float x = i * (1.f / 32);
        volatile float x1 = x + 1;
        volatile float t0 = A * x1;
        volatile float t1 = t0 - 5 * A;
        volatile float t2 = t1 * x1;
        volatile float t3 = t2 + 8 * A;
        volatile float t4 = t3 * x1;
        tab[4 * i + 0] = t4 - 4 * A;
        volatile float u0 = (A + 2) * x;
        volatile float u1 = u0 - (A + 3);
        volatile float u2 = u1 * x;
        volatile float u3 = u2 * x;
        tab[4 * i + 1] = u3 + 1;
        volatile float y = 1 - x;
        volatile float v0 = (A + 2) * y;
        volatile float v1 = v0 - (A + 3);
        volatile float v2 = v1 * y;
        volatile float v3 = v2 * y;
        tab[4 * i + 2] = v3 + 1;
        volatile float s0 = 1.f - tab[4 * i + 0];
        volatile float s1 = s0 - tab[4 * i + 1];
        tab[4 * i + 3] = s1 - tab[4 * i + 2];
    }
}

static int upload_tab()
{
    if (!g_tab_ready) {
        cubic_tab_host(g_tab_host);
        g_tab_ready = true;
    }
    return OIP_OK;
}

} // namespace pan
} // namespace oip

using namespace oip;

extern "C" int oip_pan_out_width(int n_ccd, int w, int fold_half) { return n_ccd * w - 2 * (n_ccd - 1) * fold_half; }

extern "C" void oip_cubic_tab(float *tab128)
{
    pan::upload_tab();
    memcpy(tab128, pan::g_tab_host, sizeof pan::g_tab_host);
}

static int validate_desc(const oip_pan_desc *d)
{
    if (!d) return fail(OIP_E_INVALID, "null descriptor");
    if (d->n_ccd < 1 || d->n_ccd > 8) return fail(OIP_E_INVALID, "n_ccd=%d out of range 1..8", d->n_ccd);
    if (d->w < 8 || d->w > 32760) return fail(OIP_E_INVALID, "w=%d out of range", d->w);
    if (d->fold_half < 0 || 2 * d->fold_half >= d->w) return fail(OIP_E_INVALID, "fold_half=%d invalid for w=%d", d->fold_half, d->w);
    if (d->total_rows < 0 || d->row0 < 0 || d->n_rows < 0 || d->row0 + d->n_rows > d->total_rows)
        return fail(OIP_E_INVALID, "row range [%lld,+%lld) outside strip of %lld rows", (long long)d->row0,
                    (long long)d->n_rows, (long long)d->total_rows);
    if (d->section_rows > d->row_guard) return fail(OIP_E_INVALID, "section_rows must be <= row_guard");
    for (int i = 0; i < d->n_ccd; ++i) {
        const oip_ccd_src &c = d->ccd[i];
        if (c.fmt < OIP_FMT_LE16 || c.fmt > OIP_FMT_BE16_TILES) return fail(OIP_E_INVALID, "ccd %d: unknown sample format %d", i, c.fmt);
        if (c.n_seg < 1 || c.n_seg > OIP_MAX_SEG) return fail(OIP_E_INVALID, "ccd %d: n_seg=%d", i, c.n_seg);
        if (c.fmt == OIP_FMT_BE16_TILES && (!c.d_tile_off || c.tile_cols * 8 != d->w || c.tile_lines < 1))
            return fail(OIP_E_INVALID, "ccd %d: tile layout needs d_tile_off and 8*tile_cols == w", i);
        if (c.shifted && !(std::fabs(c.dX) < 16000.0 && std::fabs(c.dY) < 16000.0))
            return fail(OIP_E_INVALID, "ccd %d: shift (%g,%g) out of range", i, c.dX, c.dY);
    }
    return OIP_OK;
}

// which CCDs may use the fast kernel (TMA tensor copies, 64/128-bit stores): alignment and format
static void fast_eligibility(const oip_pan_desc *d, bool enable, bool *fast_ccd)
{
    const bool out_ok = (((uintptr_t)d->d_out & 15) == 0) && (d->out_pitch_px % 8 == 0);
    for (int i = 0; i < d->n_ccd; ++i) {
        const oip_ccd_src &c = d->ccd[i];
        if (c.fmt == OIP_FMT_BE16_TILES) {
            // frame tiles: gathered by 4-byte cp.async, so stream base and every sub-image offset must be 4-byte
            // aligned (frames are multiples of 4 bytes long: true unless the stream starts with odd junk)
            bool ok = enable && out_ok && c.h_tile_off && c.d_tile_off && (((uintptr_t)c.d_kb & 15) == 0) && d->w >= 64 && c.n_seg == 1 &&
                      c.seg[0].base && (((uintptr_t)c.seg[0].base & 3) == 0) && c.seg[0].row0 == 0 && c.tile_lines >= 1 &&
                      c.tile_cols >= 2 * panfast::BOX_W /* a stage window spans at most two sub-image columns */ && (c.tile_cols & 1) == 0 && c.seg[0].n_rows > 0 && c.seg[0].n_rows < (1ll << 31);
            if (ok) {
                const int64_t lpf = 4ll * c.tile_lines, n_fr = (c.seg[0].n_rows + lpf - 1) / lpf;
                for (int64_t k = 0; k < n_fr * 40 && ok; ++k) {
                    if (k % 40 >= 32) continue; // MSS sub-images are not read here
                    ok = c.h_tile_off[k] < 0 || (c.h_tile_off[k] & 3) == 0;
                }
            }
            fast_ccd[i] = ok;
            continue;
        }
        const bool packed = c.fmt == OIP_FMT_PACK12 || c.fmt == OIP_FMT_PACK10;
        // packed lines: the line is addressed as 32-bit elements and unpacked in 16-sample groups
        bool ok = enable && out_ok && (c.fmt == OIP_FMT_LE16 || c.fmt == OIP_FMT_BE16 || (packed && d->w % 16 == 0)) &&
                  (((uintptr_t)c.d_kb & 15) == 0) && d->w >= 64;
        const int64_t min_pitch = packed ? (int64_t)d->w * (c.fmt == OIP_FMT_PACK12 ? 12 : 10) / 8 : 2 * (int64_t)d->w;
        for (int s = 0; s < c.n_seg && ok; ++s)
            ok = c.seg[s].base && (((uintptr_t)c.seg[s].base & 15) == 0) && (c.seg[s].pitch_bytes % 16 == 0) &&
                 c.seg[s].pitch_bytes >= min_pitch && c.seg[s].n_rows > 0;
        fast_ccd[i] = ok;
    }
}

/* host-only diagnostic: how the planner splits [row0,row0+n_rows) x out_w between the two kernels.  cover (may be
 * null) receives, per output pixel, 1 = generic kernel, 2 = fast kernel, summed if a pixel were planned twice. */
extern "C" int oip_pan_plan_coverage(const oip_pan_desc *d, int enable_fast, int fast_rows, uint8_t *cover, int64_t stats[4])
{
    int rc = validate_desc(d);
    if (rc) return rc;
    bool fast_ccd[8] = {};
    fast_eligibility(d, enable_fast != 0, fast_ccd);
    std::vector<pan::Tile> tiles;
    std::vector<panfast::FastTile> ftiles;
    int64_t cls_ctas[3] = {0, 0, 0};
    rc = pan::build_plan(d, fast_ccd, fast_rows, tiles, ftiles, cls_ctas);
    if (rc) return rc;
    int64_t px_gen = 0, px_fast = 0, n_fast = 0;
    const int64_t pitch = d->out_pitch_px;
    for (const pan::Tile &t : tiles) {
        px_gen += (int64_t)(t.x_end - t.x_begin) * t.n_rows;
        if (cover)
            for (int64_t r = 0; r < t.n_rows; ++r)
                for (int x = 0; x < t.x_end - t.x_begin; ++x) cover[(t.g0 - d->row0 + r) * pitch + t.out_x + x] += 1;
    }
    for (const panfast::FastTile &t : ftiles) {
        if (t.kind < 0) continue;
        ++n_fast;
        const int nc = t.kind == panfast::FT_REMAP ? 2 * t.half : t.half; // COPY / EDGE: half = columns
        px_fast += (int64_t)nc * t.n_rows;
        if (cover)
            for (int64_t r = 0; r < t.n_rows; ++r)
                for (int x = 0; x < nc; ++x) cover[t.out_off + r * pitch + x] += 2;
    }
    if (stats) { stats[0] = px_gen; stats[1] = px_fast; stats[2] = (int64_t)tiles.size(); stats[3] = n_fast; }
    return OIP_OK;
}

extern "C" int oip_pan_pipeline(oip_ctx *ctx, const oip_pan_desc *d)
{
    OIP_CHECK_CTX(ctx);
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!d->d_out && d->n_rows > 0) return fail(OIP_E_INVALID, "d_out is null");
    if (d->n_rows == 0) return OIP_OK;
    const int out_w = oip_pan_out_width(d->n_ccd, d->w, d->fold_half);
    if (d->out_pitch_px < out_w) return fail(OIP_E_INVALID, "out_pitch_px=%lld < out_w=%d", (long long)d->out_pitch_px, out_w);

    bool fast_ccd[8] = {};
    fast_eligibility(d, ctx->pan_fast != 0, fast_ccd);

    // ---- plan (cached on the geometry)
    pan::PlanKey key{};
    key.n_ccd = d->n_ccd; key.w = d->w; key.fold_half = d->fold_half; key.section_rows = d->section_rows;
    key.row_guard = d->row_guard; key.total_rows = d->total_rows; key.row0 = d->row0;
    key.n_rows = d->n_rows; key.out_pitch = d->out_pitch_px; key.fast_rows = ctx->pan_fast_rows;
    for (int i = 0; i < d->n_ccd; ++i) {
        key.dX[i] = d->ccd[i].dX; key.dY[i] = d->ccd[i].dY; key.shifted[i] = d->ccd[i].shifted != 0;
        key.fast[i] = fast_ccd[i]; key.n_seg[i] = d->ccd[i].n_seg; key.fmt[i] = d->ccd[i].fmt;
        if (d->ccd[i].fmt == OIP_FMT_BE16_TILES) { key.tile_cols[i] = d->ccd[i].tile_cols; key.tile_lines[i] = d->ccd[i].tile_lines; }
        for (int s = 0; s < d->ccd[i].n_seg; ++s) { key.seg_row0[i][s] = d->ccd[i].seg[s].row0; key.seg_rows[i][s] = d->ccd[i].seg[s].n_rows; }
    }
    const uint8_t *kb = reinterpret_cast<const uint8_t *>(&key);
    constexpr size_t MAX_PLANS = 96;
    oip_pan_plan *pl = nullptr;
    for (oip_pan_plan &c : ctx->pan_plans)
        if (c.key.size() == sizeof key && memcmp(c.key.data(), kb, sizeof key) == 0) { pl = &c; break; }
    if (!pl) {
        std::vector<pan::Tile> tiles;
        std::vector<panfast::FastTile> ftiles;
        int64_t cls_ctas[3] = {0, 0, 0};
        rc = pan::build_plan(d, fast_ccd, ctx->pan_fast_rows, tiles, ftiles, cls_ctas);
        if (rc) return rc;
        pan::upload_tab();
        // a free slot, else the least recently used one (its buffer may still be read by kernels in flight)
        for (oip_pan_plan &c : ctx->pan_plans)
            if (c.key.empty()) { pl = &c; break; }
        if (!pl && ctx->pan_plans.size() < MAX_PLANS) {
            ctx->pan_plans.reserve(MAX_PLANS); // pointers into the vector stay valid
            ctx->pan_plans.emplace_back();
            pl = &ctx->pan_plans.back();
        }
        if (!pl) {
            pl = &ctx->pan_plans[0];
            for (oip_pan_plan &c : ctx->pan_plans)
                if (c.last_use < pl->last_use) pl = &c;
        }
        const size_t fast_off = (1024 + tiles.size() * sizeof(pan::Tile) + 63) / 64 * 64;
        const size_t bytes = fast_off + ftiles.size() * sizeof(panfast::FastTile) + 64;
        // a kernel may still read the slot's old plan: wait for the last launch that used it (the least recently used
        // slot of a long host-buffer run finished long ago, so this does not drain the pipeline)
        if (pl->done) OIP_CUDA(cudaEventSynchronize(pl->done));
        else OIP_CUDA(cudaEventCreateWithFlags(&pl->done, cudaEventDisableTiming));
        if (bytes > pl->cap) {
            if (pl->d_plan) { OIP_CUDA(cudaFree(pl->d_plan)); pl->d_plan = nullptr; pl->cap = 0; }
            OIP_CUDA(cudaMalloc(&pl->d_plan, bytes + bytes / 4));
            pl->cap = bytes + bytes / 4;
        }
        if (bytes > pl->h_cap) {
            if (pl->h_stage) { OIP_CUDA(cudaFreeHost(pl->h_stage)); pl->h_stage = nullptr; pl->h_cap = 0; }
            OIP_CUDA(cudaHostAlloc(&pl->h_stage, bytes + bytes / 4, cudaHostAllocDefault));
            pl->h_cap = bytes + bytes / 4;
        }
        // header (weight table + the run-time (-0.0,-0.0) addend of the packed products, see mul2()), generic tiles,
        // fast warp-tiles: ONE copy from pinned staging
        uint8_t *hs = (uint8_t *)pl->h_stage;
        memset(hs, 0, 1024);
        memcpy(hs, pan::g_tab_host, 512);
        reinterpret_cast<float *>(hs)[128] = reinterpret_cast<float *>(hs)[129] = -0.0f;
        if (!tiles.empty()) memcpy(hs + 1024, tiles.data(), tiles.size() * sizeof(pan::Tile));
        if (!ftiles.empty()) memcpy(hs + fast_off, ftiles.data(), ftiles.size() * sizeof(panfast::FastTile));
        OIP_CUDA(cudaMemcpyAsync(pl->d_plan, hs, fast_off + ftiles.size() * sizeof(panfast::FastTile), cudaMemcpyHostToDevice, ctx->stream));
        pl->key.assign(kb, kb + sizeof key);
        pl->tiles = (int64_t)tiles.size();
        pl->fast_ctas = (int64_t)(ftiles.size() / panfast::WARPS);
        for (int c = 0; c < 3; ++c) pl->fast_cls[c] = cls_ctas[c];
        pl->fast_off = fast_off;
    }
    pl->last_use = ++ctx->plan_clock;
    ctx->d_plan = pl->d_plan;
    ctx->plan_tiles = pl->tiles;
    ctx->plan_fast_ctas = pl->fast_ctas;
    for (int c = 0; c < 3; ++c) ctx->plan_fast_cls[c] = pl->fast_cls[c];
    ctx->plan_fast_off = pl->fast_off;
    if (ctx->plan_tiles == 0 && ctx->plan_fast_ctas == 0) return OIP_OK;

    // the two kernels write disjoint pixels: the (small) generic launch runs on a side stream next to the fast one
    const bool fork = ctx->plan_fast_ctas > 0 && ctx->plan_tiles > 0 && ctx->side_stream;
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

    if (ctx->plan_tiles > 0) {
        // ---- generic kernel: borders, section edges, irregular map positions, packed / tiled / unaligned inputs
        pan::Params P{};
        bool bulk_ok = d->w % 8 == 0;
        for (int i = 0; i < d->n_ccd; ++i) {
            const oip_ccd_src &c = d->ccd[i];
            pan::CcdDev &o = P.ccd[i];
            o.fmt = c.fmt; o.n_seg = c.n_seg; o.kb = c.d_kb; o.dX = c.dX; o.dY = c.dY;
            o.tile_off = c.d_tile_off; o.tile_cols = c.tile_cols; o.tile_lines = c.tile_lines;
            if (c.fmt != OIP_FMT_LE16 && c.fmt != OIP_FMT_BE16) bulk_ok = false;
            for (int s = 0; s < c.n_seg; ++s) {
                if (!c.seg[s].base) return fail(OIP_E_INVALID, "ccd %d segment %d: null base", i, s);
                o.seg[s].base = (const uint8_t *)c.seg[s].base;
                o.seg[s].row0 = c.seg[s].row0; o.seg[s].n_rows = c.seg[s].n_rows; o.seg[s].pitch = c.seg[s].pitch_bytes;
                if (((uintptr_t)c.seg[s].base & 15) || (c.seg[s].pitch_bytes & 15)) bulk_ok = false;
            }
        }
        P.tab = reinterpret_cast<const float *>(ctx->d_plan);
        P.tiles = reinterpret_cast<const pan::Tile *>((const uint8_t *)ctx->d_plan + 1024);
        P.out = d->d_out; P.out_pitch = d->out_pitch_px; P.out_row0 = d->row0;
        P.err = ctx->d_err; P.w = d->w; P.n_ccd = d->n_ccd; P.bulk_ok = bulk_ok ? 1 : 0;

        const size_t smem = (size_t)pan::RING_ROWS * pan::SLOTS * 8 + 2 * (size_t)pan::STG * pan::SWC * 2;
        if (!ctx->pan_attr_set) {
            OIP_CUDA(cudaFuncSetAttribute(pan::pan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->pan_attr_set = true;
        }
        pan::pan_kernel<<<(unsigned)ctx->plan_tiles, pan::NT, smem, gstream>>>(P);
        OIP_CUDA(cudaGetLastError());
        ctx->launches++;
    }

    // ---- fast kernel: regular interior warp-tiles
    if (ctx->plan_fast_ctas > 0) {
        panfast::FastParams F;
        memset(&F, 0, sizeof F);
        for (int i = 0; i < d->n_ccd; ++i) {
            if (!fast_ccd[i]) continue;
            const oip_ccd_src &c = d->ccd[i];
            F.ccd[i].kb = c.d_kb;
            if (c.fmt == OIP_FMT_BE16_TILES) {
                panfast::FastCcd &o = F.ccd[i];
                o.swap = 1; o.tiled = 1;
                o.tile_base = (const uint8_t *)c.seg[0].base; o.tile_off = c.d_tile_off;
                o.tile_cols = c.tile_cols; o.tile_lines = c.tile_lines;
                o.n_frames = (int32_t)((c.seg[0].n_rows + 4ll * c.tile_lines - 1) / (4ll * c.tile_lines));
                o.div_lpf = panfast::fast_div_make(4u * (uint32_t)c.tile_lines);
                o.div_tl = panfast::fast_div_make((uint32_t)c.tile_lines);
                continue;
            }
            const int pbits = c.fmt == OIP_FMT_PACK12 ? 12 : (c.fmt == OIP_FMT_PACK10 ? 10 : 0);
            for (int s = 0; s < c.n_seg; ++s) {
                rc = pbits ? panfast::fast_encode_tmap_packed(&F.tmap[i * OIP_MAX_SEG + s], c.seg[s].base, d->w, pbits, c.seg[s].n_rows, c.seg[s].pitch_bytes)
                           : panfast::fast_encode_tmap(&F.tmap[i * OIP_MAX_SEG + s], c.seg[s].base, d->w, c.seg[s].n_rows, c.seg[s].pitch_bytes);
                if (rc) return rc;
            }
            F.ccd[i].swap = c.fmt == OIP_FMT_BE16;
            F.ccd[i].pbits = pbits;
        }
        F.tiles = reinterpret_cast<const panfast::FastTile *>((const uint8_t *)ctx->d_plan + ctx->plan_fast_off);
        F.out = d->d_out; F.out_pitch = d->out_pitch_px;
        F.tab = reinterpret_cast<const float *>(ctx->d_plan);
        F.w = d->w;
        F.n_stage = std::max(2, std::min(8, ctx->pan_fast_stages));
        rc = panfast::fast_launch(ctx, F, ctx->plan_fast_cls);
        if (rc) return rc;
    }
    if (fork) {
        OIP_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
        OIP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    if (pl->done) OIP_CUDA(cudaEventRecord(pl->done, ctx->stream));
    return OIP_OK;
}

extern "C" int oip_pan_check_error(oip_ctx *ctx)
{
    OIP_CHECK_CTX(ctx);
    int e = 0;
    OIP_CUDA(cudaMemcpyAsync(&e, ctx->d_err, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e) {
        OIP_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
        return fail(OIP_E_RANGE, "a kernel needed a source row that no segment supplies (halo planning error)");
    }
    return OIP_OK;
}

extern "C" int oip_pan_rows_needed(const oip_pan_desc *d, int ccd, int64_t *first, int64_t *last,
                                   int64_t *stale_first, int64_t *stale_last)
{
    int rc = validate_desc(d);
    if (rc) return rc;
    if (ccd < 0 || ccd >= d->n_ccd) return fail(OIP_E_INVALID, "ccd index");
    int64_t f = INT64_MAX, l = INT64_MIN, sf = INT64_MAX, sl = INT64_MIN;
    const bool shifted = d->ccd[ccd].shifted != 0;
    if (!shifted) {
        f = d->row0; l = d->row0 + d->n_rows;
    } else {
        std::vector<ShiftSegment> segs;
        if (!plan_shift_segments(d->total_rows, d->section_rows, d->row_guard, d->ccd[ccd].dY, segs))
            return fail(OIP_E_INVALID, "invalid section geometry");
        for (const ShiftSegment &s : segs) {
            int64_t g0 = std::max<int64_t>(s.g0, d->row0), g1 = std::min<int64_t>(s.g1, d->row0 + d->n_rows);
            if (g1 <= g0) continue;
            int64_t ta = tap_base(s.j0 + (g0 - s.g0), d->ccd[ccd].dY);
            int64_t tb = (int64_t)tap_base(s.j0 + (g1 - 1 - s.g0), d->ccd[ccd].dY) + 3;
            ta = std::max<int64_t>(ta, 0); tb = std::min<int64_t>(tb, s.hbuf - 1);
            if (tb < ta) continue;
            // fresh part
            int64_t fa = ta, fb = std::min<int64_t>(tb, s.rows_s - 1);
            if (fb >= fa) { f = std::min(f, s.sec_off + fa); l = std::max(l, s.sec_off + fb + 1); }
            int64_t sa = std::max<int64_t>(ta, s.rows_s), sb = tb;
            if (sb >= sa && s.stale_off >= 0) { sf = std::min(sf, s.stale_off + sa); sl = std::max(sl, s.stale_off + sb + 1); }
        }
    }
    if (f == INT64_MAX) { f = 0; l = 0; }
    if (sf == INT64_MAX) { sf = 0; sl = 0; }
    if (first) *first = f;
    if (last) *last = l;
    if (stale_first) *stale_first = sf;
    if (stale_last) *stale_last = sl;
    return OIP_OK;
}

// The same, as a list of disjoint row ranges: the stale rows of a partial last section come from two places of the
// previous section (just below the fresh rows, and the last rows of the 30000-row buffer), ~27000 rows apart -- a
// bounding range would make a host pipeline copy the whole section.  Ranges closer than 64 rows are merged.
extern "C" int oip_pan_row_ranges(const oip_pan_desc *d, int ccd, int64_t *ranges, int max_ranges, int *n_ranges)
{
    int rc = validate_desc(d);
    if (rc) return rc;
    if (ccd < 0 || ccd >= d->n_ccd || !ranges || !n_ranges || max_ranges < 1) return fail(OIP_E_INVALID, "oip_pan_row_ranges: bad argument");
    std::vector<std::pair<int64_t, int64_t>> iv;
    if (!d->ccd[ccd].shifted) {
        iv.push_back({d->row0, d->row0 + d->n_rows});
    } else {
        std::vector<ShiftSegment> segs;
        if (!plan_shift_segments(d->total_rows, d->section_rows, d->row_guard, d->ccd[ccd].dY, segs))
            return fail(OIP_E_INVALID, "invalid section geometry");
        for (const ShiftSegment &s : segs) {
            int64_t g0 = std::max<int64_t>(s.g0, d->row0), g1 = std::min<int64_t>(s.g1, d->row0 + d->n_rows);
            if (g1 <= g0) continue;
            int64_t ta = tap_base(s.j0 + (g0 - s.g0), d->ccd[ccd].dY);
            int64_t tb = (int64_t)tap_base(s.j0 + (g1 - 1 - s.g0), d->ccd[ccd].dY) + 3;
            ta = std::max<int64_t>(ta, 0); tb = std::min<int64_t>(tb, s.hbuf - 1);
            if (tb < ta) continue;
            const int64_t fa = ta, fb = std::min<int64_t>(tb, s.rows_s - 1);
            if (fb >= fa) iv.push_back({s.sec_off + fa, s.sec_off + fb + 1});
            const int64_t sa = std::max<int64_t>(ta, s.rows_s), sb = tb;
            if (sb >= sa && s.stale_off >= 0) iv.push_back({s.stale_off + sa, s.stale_off + sb + 1});
        }
    }
    std::sort(iv.begin(), iv.end());
    std::vector<std::pair<int64_t, int64_t>> m;
    for (const auto &r : iv) {
        if (r.second <= r.first) continue;
        if (!m.empty() && r.first <= m.back().second + 64) m.back().second = std::max(m.back().second, r.second);
        else m.push_back(r);
    }
    while ((int)m.size() > max_ranges) { // too many: merge the two closest
        size_t best = 0;
        for (size_t i = 1; i + 1 < m.size(); ++i)
            if (m[i + 1].first - m[i].second < m[best + 1].first - m[best].second) best = i;
        m[best].second = m[best + 1].second;
        m.erase(m.begin() + best + 1);
    }
    *n_ranges = (int)m.size();
    for (size_t i = 0; i < m.size(); ++i) { ranges[2 * i] = m[i].first; ranges[2 * i + 1] = m[i].second; }
    return OIP_OK;
}

extern "C" int oip_shift_cubic_u16(oip_ctx *ctx, const uint16_t *d_src, uint16_t *d_dst, int w, int64_t rows,
                                   double dX, double dY, int section_rows, int row_guard)
{
    oip_pan_desc d{};
    d.n_ccd = 1; d.w = w; d.total_rows = rows; d.row0 = 0; d.n_rows = rows; d.fold_half = 0;
    d.section_rows = section_rows; d.row_guard = row_guard;
    d.ccd[0].fmt = OIP_FMT_LE16; d.ccd[0].n_seg = 1; d.ccd[0].shifted = 1;
    d.ccd[0].seg[0] = {d_src, 0, rows, (int64_t)w * 2};
    d.ccd[0].dX = dX; d.ccd[0].dY = dY;
    d.d_out = d_dst; d.out_pitch_px = w;
    return oip_pan_pipeline(ctx, &d);
}

extern "C" int oip_stitch_concat_u16(oip_ctx *ctx, const uint16_t *const *d_ccd, int n_ccd, int w, int64_t rows,
                                     int fold_half, uint16_t *d_dst)
{
    if (n_ccd < 1 || n_ccd > 8 || !d_ccd) return fail(OIP_E_INVALID, "n_ccd=%d out of range 1..8", n_ccd);
    oip_pan_desc d{};
    d.n_ccd = n_ccd; d.w = w; d.total_rows = rows; d.row0 = 0; d.n_rows = rows; d.fold_half = fold_half;
    d.section_rows = 30000; d.row_guard = 32767;
    for (int i = 0; i < n_ccd; ++i) {
        d.ccd[i].fmt = OIP_FMT_LE16; d.ccd[i].n_seg = 1; d.ccd[i].shifted = 0;
        d.ccd[i].seg[0] = {d_ccd[i], 0, rows, (int64_t)w * 2};
    }
    d.d_out = d_dst; d.out_pitch_px = oip_pan_out_width(n_ccd, w, fold_half);
    return oip_pan_pipeline(ctx, &d);
}
