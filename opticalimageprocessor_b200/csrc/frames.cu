// frames.cu -- stage 1: AOS sync search + CRC + chained scan, IMTR re-framing, image-frame index,
// sub-image unpack.  All integer / byte work; results are bit-exact with the reference's
// sequential loops (ref aux_separator.h:256-690), whose order-dependent rules are resolved with
// parallel candidate evaluation + a short dependent walk over the (tiny) candidate table.
#include <algorithm>

#include "crc_bitslice.cuh"
#include "oip_common.cuh"

namespace oip {
namespace frames {

// =============================================================================================
// CRC-16/CCITT-FALSE split over the 32 lanes of a warp (ref CRC.h:806-834 is bit-serial).
// lane l takes bytes [l*cs, (l+1)*cs) with a zero initial remainder; the pieces are joined with
// rem = XOR_l (rem_l * x^(8*bytes_after_l) mod P)  ^  (0xFFFF * x^(8*len) mod P).
// =============================================================================================
struct CrcPlan {
    int len, cs;           // message bytes, bytes per lane
    uint16_t mul[32];      // x^(8*bytes_after_l) mod P
    uint16_t init_term;    // 0xFFFF * x^(8*len) mod P
};

static uint16_t h_mulx8(uint16_t r)
{
    for (int i = 0; i < 8; ++i) r = (uint16_t)((r & 0x8000) ? ((r << 1) ^ 0x1021) : (r << 1));
    return r;
}
static uint16_t h_gfmul(uint16_t a, uint16_t b)
{
    uint16_t r = 0;
    for (int i = 15; i >= 0; --i) {
        r = (uint16_t)((r & 0x8000) ? ((r << 1) ^ 0x1021) : (r << 1));
        if ((b >> i) & 1) r ^= a;
    }
    return r;
}
static CrcPlan make_crc_plan(int len)
{
    CrcPlan p{};
    p.len = len;
    p.cs = ((len + 31) / 32 + 3) & ~3; // multiple of 4 bytes per lane
    // xp[m] = x^(8m) mod P
    std::vector<uint16_t> xp(len + 1);
    xp[0] = 1;
    for (int m = 1; m <= len; ++m) xp[m] = h_mulx8(xp[m - 1]);
    for (int l = 0; l < 32; ++l) {
        int end = std::min(len, (l + 1) * p.cs);
        p.mul[l] = xp[len - end];
    }
    p.init_term = h_gfmul(0xFFFF, xp[len]);
    return p;
}

__device__ __forceinline__ uint32_t gfmul16(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
        r = ((r << 1) ^ ((r & 0x8000u) ? 0x1021u : 0u)) & 0xFFFFu;
        r ^= ((b >> i) & 1u) ? a : 0u;
    }
    return r;
}

// get(i) returns message byte i; every lane of the warp must call this
template <typename GetByte>
__device__ __forceinline__ uint32_t warp_crc16(const CrcPlan &P, GetByte get)
{
    const int lane = threadIdx.x & 31;
    const int b0 = lane * P.cs, b1 = min(P.len, b0 + P.cs);
    uint32_t r = 0;
    for (int i = b0; i < b1; ++i) r = crc16_byte(r, get(i));
    uint32_t c = gfmul16(r, P.mul[lane]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
    return c ^ P.init_term;
}

// =============================================================================================
// generic exclusive scan (u32), three small kernels; used on candidate tables only
// =============================================================================================
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t s_w[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? s_w[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_w[lane] = w;
    }
    __syncthreads();
    const uint32_t base = wid ? s_w[wid - 1] : 0;
    if (total) *total = s_w[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void scan_block_kernel(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *block_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(s, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void scan_add_kernel(uint32_t *out, int64_t n, const uint32_t *block_prefix)
{
    const int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
    const uint32_t add = block_prefix[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        int64_t j = i + (int64_t)k * SCAN_T;
        if (j < n) out[j] += add;
    }
}
// scratch must hold ceil(n/SCAN_B) + recursive levels u32; returns via d_total (device u32)
static int exclusive_scan_u32(oip_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *scratch,
                              uint32_t *d_total)
{
    if (n <= 0) {
        OIP_CUDA(cudaMemsetAsync(d_total, 0, 4, ctx->stream));
        return OIP_OK;
    }
    const int64_t nb = (n + SCAN_B - 1) / SCAN_B;
    scan_block_kernel<<<(unsigned)nb, SCAN_T, 0, ctx->stream>>>(in, out, n, scratch);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    if (nb == 1) {
        OIP_CUDA(cudaMemcpyAsync(d_total, scratch, 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return OIP_OK;
    }
    uint32_t *prefix = scratch + nb;
    int rc = exclusive_scan_u32(ctx, scratch, prefix, nb, prefix + nb, d_total);
    if (rc) return rc;
    scan_add_kernel<<<(unsigned)nb, SCAN_T, 0, ctx->stream>>>(out, n, prefix);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}
static size_t scan_scratch_elems(int64_t n)
{
    size_t t = 0;
    while (n > 1) {
        n = (n + SCAN_B - 1) / SCAN_B;
        t += (size_t)n * 2 + 8;
    }
    return t + 16;
}

// =============================================================================================
// AOS: sync search + the validation rules that need no CRC.  One CTA streams a 32 KiB chunk of the file straight
// from global memory (every thread has its eight 16-byte loads in flight at once), marks the sync hits in a
// shared-memory bitmap and turns the bitmap into the chunk's ordered candidate list.
// =============================================================================================
constexpr int CH = 32768;           // file bytes owned by a CTA
constexpr int AOS_T = 256;
constexpr int AOS_IT = CH / 16 / AOS_T;  // 16-byte steps per thread (8)
constexpr int AOS_WPT = CH / 32 / AOS_T; // bitmap words per thread (4)
constexpr int AOS_LIST = 1024;           // candidates per pass (a clean chunk has 32)

struct ChunkInfo {
    uint32_t slot0, count;
};

__global__ void __launch_bounds__(AOS_T, 6) aos_scan_kernel(const uint8_t *__restrict__ buf, int64_t n, uint32_t *cursor,
                                                         uint32_t cap, ChunkInfo *info, uint64_t *cand_off, int8_t *cand_st)
{
    __shared__ uint32_t s_bits[CH / 32];
    __shared__ uint16_t s_cand[AOS_LIST];
    __shared__ uint32_t s_slot0, s_total;

    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * CH;
    const int own = (int)min((int64_t)CH, n - c0);
    for (int i = tid; i < CH / 32; i += AOS_T) s_bits[i] = 0;
    __syncthreads();

    // ---- sync search "1A CF FC 1D" (ref aux_separator.h:29, :622-625); a hit must leave room for a whole frame
    //      (p + 1024 <= n), anything later can never be accepted or counted.  Filter: a 0x1A byte followed by a 0xCF
    //      byte (zero-byte trick on both, no false negatives), rare enough that a warp almost never enters the exact
    //      comparison.
    const bool vec = (((uintptr_t)buf) & 15) == 0 && c0 + CH + 4 <= n;
    constexpr int HALF = AOS_IT / 2; // two rounds of four 16-byte loads in flight: 40 registers, 6 CTAs per SM
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        uint4 q4[HALF];
        uint32_t nx[HALF];
        if (vec) {
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                const int it = round * HALF + i;
                const uint8_t *p = buf + c0 + 16 * (int64_t)(it * AOS_T + tid);
                q4[i] = ldg_nc_v4(p);
                if ((tid & 31) == 31) nx[i] = __ldg(reinterpret_cast<const uint32_t *>(p + 16)); // the next warp's first word
            }
#pragma unroll
            for (int i = 0; i < HALF; ++i) { // the word after my 16 bytes is the neighbour lane's first word
                const uint32_t v = __shfl_down_sync(0xffffffffu, q4[i].x, 1);
                if ((tid & 31) != 31) nx[i] = v;
            }
        } else { // last chunk of the file, or a buffer that is not 16-byte aligned: byte loads with bounds
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                const int it = round * HALF + i;
                const int64_t p0 = c0 + 16 * (int64_t)(it * AOS_T + tid);
                uint32_t w[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    uint32_t v = 0;
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        const int64_t q = p0 + 4 * k + bb;
                        v |= (uint32_t)(q < n ? buf[q] : 0) << (8 * bb);
                    }
                    w[k] = v;
                }
                q4[i] = make_uint4(w[0], w[1], w[2], w[3]);
                nx[i] = w[4];
            }
        }
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            const int it = round * HALF + i;
            const uint32_t wv[5] = {q4[i].x, q4[i].y, q4[i].z, q4[i].w, nx[i]};
            uint32_t z2[5], any = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint32_t x = wv[k] ^ 0xCFCFCFCFu;
                z2[k] = (x - 0x01010101u) & ~x;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = wv[k] ^ 0x1A1A1A1Au;
                any |= (x - 0x01010101u) & ~x & __funnelshift_r(z2[k], z2[k + 1], 8);
            }
            if (any & 0x80808080u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t v = __funnelshift_r(wv[k], wv[k + 1], 8 * b);
                        const int p = (it * AOS_T + tid) * 16 + k * 4 + b;
                        if (v == 0x1DFCCF1Au && p < own && c0 + p + 1024 <= n) atomicOr(&s_bits[p >> 5], 1u << (p & 31));
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- ordered list of hits (bitmap -> positions), AOS_WPT consecutive bitmap words per thread
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < AOS_WPT; ++k) cnt += __popc(s_bits[AOS_WPT * tid + k]);
    uint32_t tot;
    const uint32_t at0 = block_exclusive_scan(cnt, &tot);
    if (tid == 0) {
        s_total = tot;
        s_slot0 = tot ? atomicAdd(cursor, tot) : 0;
        info[blockIdx.x].slot0 = s_slot0;
        info[blockIdx.x].count = tot;
    }
    __syncthreads();
    const uint32_t total = s_total, slot0 = s_slot0;
    if (slot0 + total > cap) return; // table too small: the host re-runs with the exact size

    for (uint32_t base = 0; base < total; base += AOS_LIST) { // one pass unless the chunk is full of sync patterns
        if (cnt) {
            uint32_t at = at0;
            for (int k = 0; k < AOS_WPT; ++k) {
                uint32_t m = s_bits[AOS_WPT * tid + k];
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (at >= base && at < base + AOS_LIST) s_cand[at - base] = (uint16_t)((AOS_WPT * tid + k) * 32 + b);
                    ++at;
                }
            }
        }
        __syncthreads();
        // ---- ValidateAosFrame, the rules that need no CRC (ref aux_separator.h:658-677); status 2 = "the CRC decides"
        //      (aos_crc_kernel, on the candidates in file order)
        const uint32_t here = min((uint32_t)AOS_LIST, total - base);
        for (uint32_t j = tid; j < here; j += AOS_T) {
            const int mine = (int)s_cand[j];
            const uint8_t *f = buf + c0 + mine;
            const uint32_t vcid = f[5] & 0x3F;
            const uint32_t injw = ((uint32_t)f[10] << 24) | ((uint32_t)f[11] << 16) | ((uint32_t)f[12] << 8) | f[13];
            int st;
            if (injw != 0xAAAAAAAAu && injw != 0u) st = -1;          // :675
            else if (injw == 0xAAAAAAAAu && vcid == 0x3F) st = 0;   // :676
            else st = 2;                                            // :679-686
            cand_off[slot0 + base + j] = (uint64_t)(c0 + mine);
            cand_st[slot0 + base + j] = (int8_t)st;
        }
        __syncthreads();
    }
}

// =============================================================================================
// AOS, fused single pass (round 2): sync check + header rules + CRC of 32 consecutive frames per warp, every file byte
// read at most once.
//
// The reference scan (ref aux_separator.h:421-461) jumps from an accepted frame straight to its end, so the bytes INSIDE
// an accepted frame are never searched.  A downlink is a run of back-to-back 1024-byte frames: with the phase of that
// cadence known (first sync word of the file), a warp takes the 32 slots of a 32 KiB group and
//   * checks the sync word at every slot start; if one is missing the cadence is broken here and the group is searched
//     byte by byte like aos_scan_kernel does (slow path, exact for any input);
//   * evaluates ValidateAosFrame for all 32 slots: header rules from the first words, the 32 CRCs bit-sliced with
//     word-interleaved pieces straight from global memory (the 128-byte LDPC tail of a frame is not even read);
//   * searches the interior of every slot that is NOT valid (empty, bad inject word, bad CRC): the reference advances by
//     4 bytes there and goes on searching, so false sync words inside such a frame are candidates.
// The candidates go to the same table the general kernels use (order / runstart / walk / emit below).  One case is left
// to the general path: a VALID slot that the walk finds shadowed by an overlapping accepted frame -- its interior was
// not searched although the reference's scan continues inside it.  aos_walk_kernel flags it and the host repeats the
// call with the exhaustive search (two valid frames that overlap need a CRC collision: never seen, still exact).
// =============================================================================================
constexpr int GRP = 32 * 1024;      // bytes per group: 32 slots
constexpr int AOS_F_WARPS = 4;

__global__ void aos_phase_kernel(const uint8_t *__restrict__ buf, int64_t n, uint32_t *phase)
{
    // first sync word that leaves room for a frame, within the first 64 KiB (else: no cadence assumed, phase 0);
    // 64 CTAs x 256 threads x 4 positions
    const int64_t lim = min((int64_t)65536, n - 1023);
    const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    uint32_t best = 0xFFFFFFFFu;
    if (p0 < lim) {
        uint32_t b[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) b[i] = p0 + i < n ? buf[p0 + i] : 0u;
#pragma unroll
        for (int i = 3; i >= 0; --i)
            if (p0 + i < lim && b[i] == 0x1A && b[i + 1] == 0xCF && b[i + 2] == 0xFC && b[i + 3] == 0x1D) best = (uint32_t)(p0 + i);
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0xFFFFFFFFu) atomicMin(phase, best);
}

// sync hits among the file positions [row + lo, row + hi) of one 1024-byte row, with p + 1024 <= n: lane l owns the
// positions 32l .. 32l+31 and returns their hit mask
__device__ __forceinline__ uint32_t aos_row_hits(const uint8_t *__restrict__ buf, int64_t n, int64_t row, int lo, int hi)
{
    const int lane = threadIdx.x & 31;
    const int64_t p0 = row + 32 * lane;
    uint32_t w[9];
    if (((((uintptr_t)buf) + p0) & 3) == 0 && p0 + 36 <= n) {
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<const uint32_t *>(buf + p0) + k);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t q = p0 + 4 * k + b;
                v |= (uint32_t)(q >= 0 && q < n ? buf[q] : 0) << (8 * b);
            }
            w[k] = v;
        }
    }
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t x = w[k] ^ 0x1A1A1A1Au;
        any |= (x - 0x01010101u) & ~x;
    }
    uint32_t mask = 0;
    if (any & 0x80808080u) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int pos = 32 * lane + 4 * k + b;
                if (__funnelshift_r(w[k], w[k + 1], 8 * b) == 0x1DFCCF1Au && pos >= lo && pos < hi && row + pos + 1024 <= n)
                    mask |= 1u << (4 * k + b);
            }
    }
    return mask;
}
// ValidateAosFrame without the CRC (ref aux_separator.h:658-677): -1 invalid, 0 empty, 2 = "the CRC decides"
__device__ __forceinline__ int aos_rules(const uint8_t *f)
{
    const uint32_t vcid = f[5] & 0x3F;
    const uint32_t injw = ((uint32_t)f[10] << 24) | ((uint32_t)f[11] << 16) | ((uint32_t)f[12] << 8) | f[13];
    if (injw != 0xAAAAAAAAu && injw != 0u) return -1;
    if (injw == 0xAAAAAAAAu && vcid == 0x3F) return 0;
    return 2;
}
// emit this lane's hits of one row in file order at table position `at` (warp-wide exclusive prefix); returns the count
__device__ __forceinline__ uint32_t aos_emit_row(const uint8_t *__restrict__ buf, int64_t row, uint32_t mask, uint32_t at, uint32_t cap,
                                                 uint64_t *cand_off, int8_t *cand_st)
{
    const int lane = threadIdx.x & 31;
    const uint32_t c = __popc(mask);
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, x, 31);
    uint32_t k = at + x - c;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const int64_t p = row + 32 * lane + b;
        if (k < cap) {
            cand_off[k] = (uint64_t)p;
            cand_st[k] = (int8_t)aos_rules(buf + p);
        }
        ++k;
    }
    return total;
}

__global__ void __launch_bounds__(AOS_F_WARPS * 32) aos_fused_kernel(const uint8_t *__restrict__ buf, int64_t n, const uint32_t *__restrict__ phase_ptr,
                                                                      int64_t n_groups, uint32_t *cursor, uint32_t cap, ChunkInfo *info,
                                                                      uint64_t *cand_off, int8_t *cand_st, uint32_t *n_irregular)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * AOS_F_WARPS + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const uint32_t ph = *phase_ptr;
    const int64_t phase = ph == 0xFFFFFFFFu ? 0 : (int64_t)ph;
    const int64_t g0 = phase + g * GRP;                       // first byte of the group = start of slot 0
    if (g0 >= n && g > 0) {                                   // (the grid is sized for phase 0)
        if (lane == 0) { info[g].slot0 = 0; info[g].count = 0; }
        return;
    }
    bool fast = g0 + GRP + 4 <= n;                            // 32 whole slots (and the word after them is readable)
    uint32_t nonvalid = 0;
    int st = 0;
    if (fast) {
        // ---- 32 slots: sync word, header rules, bit-sliced CRC over message + stored CRC (frame bytes 4..895: the
        //      remainder is zero exactly when the stored CRC matches; the 896-byte span starts AT the frame)
        const uint8_t *A = buf + g0 + 4 * lane;
        const uint32_t sh = (uint32_t)((uintptr_t)A & 3u);
        const uint32_t *W = reinterpret_cast<const uint32_t *>(A - sh);
        uint32_t m_sync = 0, m_a = 0, m_b = 0;                 // per-lane predicates over the 32 frames (bit q = frame q)
        uint32_t P[16];
        bitslice::warp_crc32frames_interleaved(
            [&](int j, uint32_t(&T)[32]) {
                const uint32_t *w = W + 32 * j;
                if (sh) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) T[q] = __funnelshift_r(__ldg(w + 256 * q), __ldg(w + 256 * q + 1), 8u * sh);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) T[q] = __ldg(w + 256 * q);
                }
                if (j == 0) {
                    // lane 0 holds word 0 (sync), lane 1 word 1 (byte 5 = VCID), lanes 2 / 3 words 2 / 3 (bytes 10..13 = inject word)
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const uint32_t t = T[q];
                        const uint32_t hi16 = t >> 16, lo16 = t & 0xFFFFu;
                        bool s, a, b;
                        if (lane == 0) { s = t == 0x1DFCCF1Au; a = b = false; }
                        else if (lane == 1) { s = ((t >> 8) & 0x3Fu) == 0x3Fu; a = b = false; }
                        else if (lane == 2) { s = false; a = hi16 == 0u; b = hi16 == 0xAAAAu; }
                        else { s = false; a = lo16 == 0u; b = lo16 == 0xAAAAu; }
                        m_sync |= (uint32_t)s << q; m_a |= (uint32_t)a << q; m_b |= (uint32_t)b << q;
                    }
                }
            },
            bitslice::SPAN - 892, P);
        const uint32_t sync_all = __shfl_sync(0xffffffffu, m_sync, 0);
        if (sync_all != 0xFFFFFFFFu) {
            fast = false;                                       // a slot start without sync word: cadence broken in this group
        } else {
            const uint32_t vc3f = __shfl_sync(0xffffffffu, m_sync, 1);
            const uint32_t inj0 = __shfl_sync(0xffffffffu, m_a, 2) & __shfl_sync(0xffffffffu, m_a, 3);
            const uint32_t injA = __shfl_sync(0xffffffffu, m_b, 2) & __shfl_sync(0xffffffffu, m_b, 3);
            const bool invalid = !(((inj0 | injA) >> lane) & 1u);                      // :675
            const bool empty = !invalid && ((injA >> lane) & 1u) && ((vc3f >> lane) & 1u); // :676
            const bool crc_ok = (bitslice::unslice(P, lane) ^ bitslice::init_term(892)) == 0u;
            st = invalid ? -1 : (empty ? 0 : (crc_ok ? 1 : -1));                        // :679-686
            nonvalid = __ballot_sync(0xffffffffu, st != 1);
        }
    }
    if (fast) {
        // ---- interiors of the slots that are not valid (the scan goes on at slot + 4)
        uint32_t extra = 0, has_hits = 0;
        for (uint32_t m = nonvalid; m; m &= m - 1) {
            const int q = __ffs(m) - 1;
            const uint32_t hm = aos_row_hits(buf, n, g0 + 1024 * q, 4, 1024);
            const uint32_t c = __reduce_add_sync(0xffffffffu, __popc(hm));
            extra += c;
            if (c) has_hits |= 1u << q;
        }
        const uint32_t total = 32u + extra;
        uint32_t slot0 = 0;
        if (lane == 0) {
            slot0 = atomicAdd(cursor, total);
            info[g].slot0 = slot0;
            info[g].count = total;
        }
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if ((uint64_t)slot0 + total > cap) return;             // table too small: the host re-runs with the exact size
        if (!has_hits) {
            cand_off[slot0 + lane] = (uint64_t)(g0 + 1024 * lane);
            cand_st[slot0 + lane] = (int8_t)st;
        } else {
            if (lane == 0) atomicAdd(n_irregular, 1u);         // candidates that still wait for their CRC (aos_crc_kernel)
            uint32_t at = slot0;
            for (int q = 0; q < 32; ++q) {
                const int sq = __shfl_sync(0xffffffffu, st, q);
                if (lane == 0) { cand_off[at] = (uint64_t)(g0 + 1024 * q); cand_st[at] = (int8_t)sq; }
                ++at;
                if ((has_hits >> q) & 1u) at += aos_emit_row(buf, g0 + 1024 * q, aos_row_hits(buf, n, g0 + 1024 * q, 4, 1024), at, cap, cand_off, cand_st);
            }
        }
        return;
    }
    // ---- slow path: every position of the group (the last, partial group; a group where the cadence is broken; a file
    //      without cadence).  Group 0 also owns the bytes in front of the phase (none of them starts a sync word).
    const int64_t lo_byte = g == 0 ? 0 : g0, hi_byte = min(n, g0 + GRP);
    uint32_t total = 0;
    for (int64_t row = lo_byte; row < hi_byte; row += 1024)
        total += __reduce_add_sync(0xffffffffu, __popc(aos_row_hits(buf, n, row, 0, (int)min((int64_t)1024, hi_byte - row))));
    uint32_t slot0 = 0;
    if (lane == 0) {
        slot0 = total ? atomicAdd(cursor, total) : 0;
        info[g].slot0 = slot0;
        info[g].count = total;
    }
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (!total || (uint64_t)slot0 + total > cap) return;
    if (lane == 0) atomicAdd(n_irregular, 1u);
    uint32_t at = slot0;
    for (int64_t row = lo_byte; row < hi_byte; row += 1024)
        at += aos_emit_row(buf, row, aos_row_hits(buf, n, row, 0, (int)min((int64_t)1024, hi_byte - row)), at, cap, cand_off, cand_st);
}

// ---- CRC of the candidates that need one (ref aux_separator.h:679-686), on the file-ordered table: a warp takes 32
// consecutive candidates.  32 frames back to back (the normal case) are bit-sliced straight from global memory with
// word-interleaved pieces: lane l reads word l of each 128-byte block of the 896-byte span that ends with the last
// message byte (starts 2 bytes before the frame), so every warp load is one contiguous 128-byte line and every byte is
// read once.  Anything else (false sync, end of file) goes one candidate at a time.
constexpr int AOS_CRC_WARPS = 4;
__global__ void __launch_bounds__(AOS_CRC_WARPS * 32) aos_crc_kernel(const uint8_t *__restrict__ buf, const __grid_constant__ CrcPlan crc,
                                                                      const uint64_t *__restrict__ off, int8_t *st_io,
                                                                      const uint32_t *__restrict__ m_ptr, const uint32_t *__restrict__ skip_unless)
{
    if (skip_unless && *skip_unless == 0u) return; // after the fused kernel: no candidate is waiting for a CRC
    const int lane = threadIdx.x & 31;
    const int64_t g0 = (((int64_t)blockIdx.x * AOS_CRC_WARPS) + (threadIdx.x >> 5)) * 32;
    const int64_t m = (int64_t)*m_ptr;
    if (g0 >= m) return;
    const int cnt = (int)min((int64_t)32, m - g0);
    const uint64_t mine = lane < cnt ? off[g0 + lane] : 0ull;
    int st = lane < cnt ? (int)st_io[g0 + lane] : 0;
    if (!__any_sync(0xffffffffu, st == 2)) return;
    const uint8_t *f = buf + mine;
    const uint32_t want = st == 2 ? (((uint32_t)f[894] << 8) | f[895]) : 0u;
    const uint64_t first = __shfl_sync(0xffffffffu, mine, 0);
    const bool uniform = __all_sync(0xffffffffu, cnt == 32 && mine == first + 1024ull * (uint64_t)lane);
    if (uniform) {
        // The remainder is taken over message + stored CRC (892 bytes, frame bytes 4..895): it is zero exactly when the
        // stored CRC matches (no reflection, no final xor).  That span of 896 bytes starts AT the frame, so the loads
        // are word aligned whenever the frame is, and lane 0 drops the 4 sync bytes in front.
        const uint8_t *A = buf + first + 4 * lane;              // span start of frame 0 + this lane's word
        const uint32_t sh = (uint32_t)((uintptr_t)A & 3u);
        const uint32_t *W = reinterpret_cast<const uint32_t *>(A - sh);
        uint32_t P[16];
        bitslice::warp_crc32frames_interleaved(
            [&](int j, uint32_t(&T)[32]) {
                const uint32_t *w = W + 32 * j;
                if (sh) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) T[q] = __funnelshift_r(__ldg(w + 256 * q), __ldg(w + 256 * q + 1), 8u * sh);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) T[q] = __ldg(w + 256 * q);
                }
            },
            bitslice::SPAN - 892, P);
        if (st == 2) st = (bitslice::unslice(P, lane) ^ bitslice::init_term(892)) == 0u ? 1 : -1;
    } else {
        for (int k = 0; k < cnt; ++k) {
            if (__shfl_sync(0xffffffffu, st, k) != 2) continue;
            const uint8_t *p = buf + __shfl_sync(0xffffffffu, mine, k) + 4;
            const uint32_t c = warp_crc16(crc, [&](int i) { return (uint32_t)p[i]; });
            if (lane == k) st = c == want ? 1 : -1;
        }
    }
    if (lane < cnt) st_io[g0 + lane] = (int8_t)st;
}

// candidates in file order: chunk c's run goes to [base[c], base[c]+count)
__global__ void aos_order_kernel(const ChunkInfo *info, const uint32_t *base, int64_t n_chunks, uint32_t cap, const uint64_t *cand_off,
                                 const int8_t *cand_st, uint64_t *ord_off, int8_t *ord_st)
{
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    const ChunkInfo ci = info[c];
    const uint32_t b = base[c];
    if ((uint64_t)ci.slot0 + ci.count > cap || (uint64_t)b + ci.count > cap) return; // table overflow: the host re-runs with the exact size
    for (uint32_t j = lane; j < ci.count; j += 32) {
        ord_off[b + j] = cand_off[ci.slot0 + j];
        ord_st[b + j] = cand_st[ci.slot0 + j];
    }
}

// the candidate count the later kernels see never exceeds the table (an overflowing call is repeated by the host)
__global__ void aos_clamp_kernel(uint32_t *total, uint32_t cap)
{
    if (*total > cap) *total = cap;
}

// Byte-range shard of a file: the scan of this shard starts at buffer offset carry_in (the previous shard's
// last accepted frame ends there) and owns the candidates that START before `own`; the bytes after `own` are the halo that
// lets the frames starting near the end be validated.  Candidates in front of carry_in get status -2 ("not there"),
// the table is cut at the first candidate >= own.
__global__ void aos_shard_trim_kernel(const uint64_t *off, int8_t *st, uint32_t *m_ptr, uint64_t own, uint64_t carry_in, uint32_t *m_own)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*m_ptr) return;
    if (off[i] < carry_in) st[i] = -2;
    if (off[i] < own) atomicMax(m_own, (uint32_t)(i + 1));
}

// A candidate with no VALID candidate in the preceding 1023 bytes cannot lie inside an accepted frame, so the sequential
// scan (ref aux_separator.h:421-461) always visits it: it starts a run that is resolved on its own.  A valid one is
// accepted there; a rejected one (empty / bad inject word / bad CRC) is counted and changes nothing for what follows.
// (Round 1 started runs at valid candidates only: a long stretch of fill frames or of corrupted frames -- or a file without
// a single valid frame -- was then walked by ONE thread.  ADVICE r1.)
__global__ void aos_runstart_kernel(const uint64_t *off, const int8_t *st, const uint32_t *m_ptr, uint8_t *rs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = (int64_t)*m_ptr;
    if (i >= m) return;
    uint8_t r = 1;
    for (int64_t j = i - 1; j >= 0 && off[i] - off[j] < 1024; --j)
        if (st[j] == 1) { r = 0; break; }
    rs[i] = r;
}

// each run start replays the reference's skip rules up to the next run start:
// accepted frame -> next search position = off+1024; rejected candidate -> off+4 (:440-441,:456-457)
__global__ void aos_walk_kernel(const uint64_t *off, const int8_t *st, const uint8_t *rs, const uint32_t *m_ptr, uint32_t *acc,
                                unsigned long long *counters, uint32_t *shadowed_valid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = (int64_t)*m_ptr;
    uint32_t n_inv = 0, n_emp = 0, n_val = 0;
    if (i < m && rs[i]) {
        uint64_t next_free = 0; // first byte the scan may look at
        for (int64_t j = i; j < m && (j == i || !rs[j]); ++j) {
            if (off[j] < next_free) { // inside an accepted frame: never seen
                acc[j] = 0;
                // a VALID frame shadowed by an overlapping accepted one: the scan continues INSIDE it, where the fused
                // kernel did not search (see aos_fused_kernel) -> the host repeats the call with the exhaustive search
                if (st[j] == 1) *shadowed_valid = 1u;
                continue;
            }
            if (st[j] == 1) {
                acc[j] = 1;
                n_val++;
                next_free = off[j] + 1024;
            } else {
                acc[j] = 0;
                if (st[j] == -1) n_inv++; else if (st[j] == 0) n_emp++; // (-2: in front of a shard's scan start, not there)
            }
        }
    }
    // in a clean downlink every frame is its own run start: the counts are added up per CTA first (one atomic per warp
    // on the same three addresses -- 83 000 of them for a 900 MB file -- serialised in L2: 41 us for a kernel that
    // moves 27 MB)
    __shared__ uint32_t s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    n_val = __reduce_add_sync(0xffffffffu, n_val);
    n_inv = __reduce_add_sync(0xffffffffu, n_inv);
    n_emp = __reduce_add_sync(0xffffffffu, n_emp);
    if ((threadIdx.x & 31) == 0) {
        if (n_val) atomicAdd(&s_cnt[0], n_val);
        if (n_inv) atomicAdd(&s_cnt[1], n_inv);
        if (n_emp) atomicAdd(&s_cnt[2], n_emp);
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void aos_emit_kernel(const uint64_t *off, const uint32_t *acc, const uint32_t *rank, const uint32_t *m_ptr,
                                uint64_t *payload_off, uint64_t cap, unsigned long long *last_end)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*m_ptr || !acc[i]) return;
    if (payload_off && rank[i] < cap) payload_off[rank[i]] = off[i] + 14; // AOS_DATA_OFF :44
    if (i + 1 == (int64_t)*m_ptr || !acc[i + 1]) atomicMax(last_end, (unsigned long long)(off[i] + 1024)); // where the scan goes on (a shard's carry-out)
}

// =============================================================================================
// IMTR: fixed 882-byte cadence over the virtual concatenation of the 880-byte payloads
// =============================================================================================
// An 882-byte frame of the virtual stream spans 2 payloads (3 when only 1 byte of it lies in the first): three
// contiguous source runs.  Frame position q lives at p0+q for q < l0, at p1+(q-l0) for q < l0+880, else at
// p2+(q-l0-880).
struct FrameSegs {
    const uint8_t *p0, *p1, *p2;
    int l0;
};
__device__ __forceinline__ FrameSegs frame_segs(const uint8_t *buf, const uint64_t *poff, int64_t n_payload, int64_t s0)
{
    const int64_t i = s0 / 880;
    const int o = (int)(s0 - i * 880);
    FrameSegs S;
    S.l0 = 880 - o;
    S.p0 = buf + poff[i] + o;
    S.p1 = i + 1 < n_payload ? buf + poff[i + 1] : S.p0;
    S.p2 = i + 2 < n_payload ? buf + poff[i + 2] : S.p1; // only touched when l0 < 2
    return S;
}
__device__ __forceinline__ const uint8_t *seg_ptr(const FrameSegs &S, int q)
{
    return q < S.l0 ? S.p0 + q : (q < S.l0 + 880 ? S.p1 + (q - S.l0) : S.p2 + (q - S.l0 - 880));
}
// frame bytes q..q+3 as a little-endian word: two aligned 32-bit loads + funnel shift, taken from the run that holds
// byte q.  Branch-free, so that all the loads of a lane are in flight together; a word that straddles a run boundary
// (<= 2 per frame) comes out wrong in its trailing bytes and is redone by seg_word_bytes.  The aligned loads touch at
// most 6 bytes past the run, which are bytes of the same 1024-byte AOS frame (CRC / LDPC field).
__device__ __forceinline__ uint32_t seg_word(const FrameSegs &S, int q)
{
    const uint8_t *a = seg_ptr(S, q);
    const uint32_t sh = (uint32_t)((uintptr_t)a & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a - sh);
    const uint32_t lo = __ldg(w), hi = __ldg(w + (sh ? 1 : 0));
    return __funnelshift_r(lo, hi, 8u * sh);
}
__device__ __forceinline__ uint32_t seg_word_bytes(const FrameSegs &S, int q, int n_bytes)
{
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (b < n_bytes) v |= (uint32_t)__ldg(seg_ptr(S, q + b)) << (8 * b);
    return v;
}
// does the word at frame position q (4 bytes) straddle one of the two run boundaries?
__device__ __forceinline__ bool seg_straddles(const FrameSegs &S, int q)
{
    return (q < S.l0 && q + 3 >= S.l0) || (q < S.l0 + 880 && q + 3 >= S.l0 + 880);
}

// (Round 2 measured a warp-private form -- every warp gathers and validates its own 32 frames, 28.8 KB of shared memory per warp,
// 7 warps per SM, loads of 4 frames in flight: 0.9 ms per 810 MB file against 0.69 ms for this one; too few warps to
// hide the latency of the gather, 500 instructions per frame.  Kept: this form.)
// A CTA validates 32 consecutive frames: its 4 warps gather 8 frames each into shared memory (896-byte slots), then
// warp 0 runs ValidateImtrFrame for all of them, lane f = frame f, with the CRCs of the 32 frames bit-sliced.
constexpr int IMTR_T = 128, IMTR_BATCH = 32, IMTR_SLOT = 896, IMTR_FRONT = 32;
__global__ void __launch_bounds__(IMTR_T) imtr_validate_kernel(const uint8_t *__restrict__ buf, const uint64_t *__restrict__ poff,
                                                               int64_t n_payload, int64_t n_frames, uint8_t *status,
                                                               uint32_t *seq, uint8_t *chid, uint32_t *valid,
                                                               unsigned long long *n_bad, uint8_t *imdt_spec, int skip)
{
    __shared__ __align__(16) uint32_t s_w[(IMTR_FRONT + IMTR_BATCH * IMTR_SLOT + 32) / 4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * IMTR_BATCH;
    // payload offsets of the whole batch in ONE round trip (every warp keeps its own copy in two registers per lane): the
    // 32 frames span payloads i0 .. i0+33 (+2 for the run tails), handed out by shuffles -- no dependent load per frame
    const int64_t s_first = (int64_t)skip + f0 * 882; // (a shard's first frame starts `skip` bytes into its first payload)
    const int64_t i0 = s_first / 880;
    const int o0 = (int)(s_first - i0 * 880);
    const unsigned long long pl0 = poff[min(i0 + lane, n_payload - 1)], pl1 = poff[min(i0 + 32 + lane, n_payload - 1)];
    auto payload = [&](int k) -> const uint8_t * { // payload i0 + k (the index is clamped to the last one like frame_segs does)
        const unsigned long long a = __shfl_sync(0xffffffffu, pl0, k & 31), b2 = __shfl_sync(0xffffffffu, pl1, k & 31);
        return buf + (k < 32 ? a : b2);
    };
    for (int q = wid; q < IMTR_BATCH; q += IMTR_T / 32) {
        if (f0 + q >= n_frames) break;
        uint32_t *fw = s_w + (IMTR_FRONT + q * IMTR_SLOT) / 4;
        FrameSegs S;
        {
            const int t = o0 + 882 * q, k = t / 880, o = t - 880 * k;   // frame q starts o bytes into payload i0 + k
            S.l0 = 880 - o;
            S.p0 = payload(k) + o;
            S.p1 = i0 + k + 1 < n_payload ? payload(k + 1) : S.p0;
            S.p2 = i0 + k + 2 < n_payload ? payload(k + 2) : S.p1;
        }
        // (splitting this loop into a load pass over the warp's 8 frames and a fix-up / output pass was measured: slower)
        // 220 whole words + 2 bytes.  ALL loads of the frame are issued before the first use of a loaded value (the raw
        // word pairs and the fix-up bytes stay in registers; the funnel shifts come afterwards): one memory round trip per
        // frame.  (Round 1 shifted each word right after its two loads and the compiler kept that order: seven dependent
        // round trips per frame -- the critical path of the CTA, ncu: long-scoreboard stalls on every SHF.)
        uint32_t lo[7], hi[7], shv[7], fxb[4];
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            const uint8_t *a = seg_ptr(S, 4 * min(lane + 32 * u, 219));
            const uint32_t sh = (uint32_t)((uintptr_t)a & 3u);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(a - sh);
            lo[u] = __ldg(w);
            hi[u] = __ldg(w + (sh ? 1 : 0));
            shv[u] = 8u * sh;
        }
        {   // lanes 0 / 1: the word across the first / second run boundary; lane 2: the last two bytes (others: a harmless reload)
            const int bnd = lane == 0 ? S.l0 : (lane == 1 ? S.l0 + 880 : 880);
            const int qb = min(bnd & ~3, 880);
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) fxb[bb] = (uint32_t)__ldg(seg_ptr(S, min(qb + bb, 881)));
        }
#pragma unroll
        for (int u = 0; u < 7; ++u)
            if (lane + 32 * u < 220) fw[lane + 32 * u] = __funnelshift_r(lo[u], hi[u], shv[u]);
        __syncwarp();
        {
            const uint32_t fx = fxb[0] | (fxb[1] << 8) | (fxb[2] << 16) | (fxb[3] << 24);
            if (lane < 2) {
                const int b = lane == 0 ? S.l0 : S.l0 + 880;
                if ((b & 3) && b < 880) fw[b >> 2] = fx;
            } else if (lane == 2) {
                fw[220] = fx & 0xFFFFu;
            }
        }
        if (imdt_spec) {
            // speculative output: in a clean downlink every frame is valid and frame f's 866 payload bytes (frame bytes
            // 10..875, IMTR_IMGDATA_OFF :72) land at f * 866 -- written here while the frame is in shared memory;
            // imtr_copy_kernel then has nothing to do.  Anything else (a rejected frame, a sequence restart) is put
            // right by imtr_copy_kernel, which rewrites the whole output from the source.
            __syncwarp();
            uint8_t *d = imdt_spec + (uint64_t)(f0 + q) * 866;
            const uint8_t *fr = reinterpret_cast<const uint8_t *>(fw);
            const int head = (int)((4u - (uint32_t)((uintptr_t)d & 3u)) & 3u);
            if (lane < head) d[lane] = fr[10 + lane];
            const int nw = (866 - head) >> 2, s0 = 10 + head;
            const uint32_t sh = 8u * (uint32_t)(s0 & 3);
            uint32_t *dw = reinterpret_cast<uint32_t *>(d + head);
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int k = lane + 32 * u;
                if (k < nw) dw[k] = __funnelshift_r(fw[(s0 >> 2) + k], fw[(s0 >> 2) + k + 1], sh);
            }
            const int t0 = head + 4 * nw;
            if (lane < 866 - t0) d[t0 + lane] = fr[10 + t0 + lane];
        }
    }
    __syncthreads();
    // ---- CRCs of the 32 frames, bit-sliced (lane l = the 28-byte piece l of every frame, bit f of a register = frame f).
    //      The seven words of a piece are split over the FOUR warps (0-1, 2-3, 4-5, 6): each warp runs its words from a
    //      zero remainder and advances the result to the end of the piece (x^(32 * words after it), a fixed XOR
    //      network); the partial remainders are added in shared memory.  (Round 1: one warp ran all seven words while
    //      the other three had nothing left to do -- the validating warp was the critical path of the CTA.)
    // :577-583 CRC over bytes 0..875: the 896-byte span ends with byte 875, i.e. starts 20 bytes before the frame
    // (whatever the previous slot left there: lane 0 drops it).  Slots and pieces are word aligned.
    __shared__ uint32_t s_part[4][16][32];
    uint32_t P[16];
    {
        const uint32_t B = smem_u32(s_w) + (uint32_t)(IMTR_FRONT - (bitslice::SPAN - 876) + 28 * lane);
        const int clear_bit = lane == 0 ? 8 * (bitslice::SPAN - 876) : -1; // lane 0: the 20 bytes in front of the frame
        const int jlo = 2 * wid, jhi = min(2 * wid + 2, 7);
#pragma unroll
        for (int i = 0; i < 16; ++i) P[i] = 0u;
#pragma unroll 1
        for (int j = jlo; j < jhi; ++j) {
            uint32_t T[32];
            const uint32_t a = B + 4u * (uint32_t)j;
#pragma unroll
            for (int q = 0; q < 32; ++q) T[q] = lds_u32(a + (uint32_t)(IMTR_SLOT * q));
            bitslice::transpose32(T);
            bitslice::lfsr_word(P, T, clear_bit - 32 * j);
        }
        if (lane == 0 && 32 * jhi <= 8 * (bitslice::SPAN - 876)) { // only bytes in front of the frame so far
#pragma unroll
            for (int i = 0; i < 16; ++i) P[i] = 0u;
        }
        uint32_t Q[16];
        if (wid == 0) bitslice::mul_xpow<160>(P, Q);
        else if (wid == 1) bitslice::mul_xpow<96>(P, Q);
        else if (wid == 2) bitslice::mul_xpow<32>(P, Q);
        else {
#pragma unroll
            for (int i = 0; i < 16; ++i) Q[i] = P[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s_part[wid][i][lane] = Q[i];
    }
    __syncthreads();
    if (wid != (int)(blockIdx.x & 3)) return; // rotate the validating warp over the SM's four schedulers
#pragma unroll
    for (int i = 0; i < 16; ++i) P[i] = s_part[0][i][lane] ^ s_part[1][i][lane] ^ s_part[2][i][lane] ^ s_part[3][i][lane];
    {
        auto x = [](uint32_t v, int sd, int) { return __shfl_xor_sync(0xffffffffu, v, sd); };
        bitslice::join_level<1>(P, lane, x);
        bitslice::join_level<2>(P, lane, x);
        bitslice::join_level<4>(P, lane, x);
        bitslice::join_level<8>(P, lane, x);
        bitslice::join_level<16>(P, lane, x);
    }
    // ValidateImtrFrame, checks in the reference's order (ref aux_separator.h:558-590)
    const uint32_t *fw = s_w + (IMTR_FRONT + lane * IMTR_SLOT) / 4;
    const uint8_t *fr = reinterpret_cast<const uint8_t *>(fw);
    int st = 0;
    if (fw[0] != 0x1FCE5449u) st = 1;                                                                // :559  49 54 CE 1F
    else if (!(fr[878] == 0x2E && fr[879] == 0xE9 && fr[880] == 0xC8 && fr[881] == 0xFD)) st = 2;     // :563
    else if (fr[9] != 0x22) st = 3;                                                                  // :572
    const uint32_t want = ((uint32_t)fr[876] << 8) | fr[877];
    if (st == 0 && (bitslice::unslice(P, lane) ^ bitslice::init_term(876)) != want) st = 4;
    const int64_t f = f0 + lane;
    if (f < n_frames) {
        status[f] = (uint8_t)st;
        seq[f] = __byte_perm(fw[1], 0u, 0x0123);                                                     // :568-569 BE u32 at 4
        chid[f] = fr[8];
        valid[f] = st == 0;
        if (st) atomicAdd(&n_bad[st - 1], 1ull); // rejected frames are rare: counted here, not on the host
    }
}

// ---- run-based form of the same kernel (the default; option imtr_runs = 0 selects the per-frame gather above) ----------
// The 32 frames of a CTA are 28224 consecutive bytes of the payload stream, i.e. pieces of at most 34 consecutive payload
// RUNS of 880 bytes.  880 and 32 * 882 are multiples of four, so when the runs are laid down back to back in shared
// memory -- run i0 + k at word 220 k -- every run is a whole number of destination words and its source is re-aligned by ONE
// funnel shift per word with a shift that is constant over the run: no per-word run selection, no boundary fix-ups.  The
// frames then start at byte o0 + 882 q of that image (any alignment when a shard starts `skip` bytes into its first payload):
// the CRC pieces and the payload words are read with two aligned LDS + a funnel shift whose amount depends on the parity of
// q only.  Per frame ~45 instructions of gather instead of ~300.
constexpr int IMTR_RUNS = 34, IMTR_RUN_WORDS = 220, IMTR_DEPTH = 3;
__global__ void __launch_bounds__(IMTR_T, 6) imtr_validate_runs_kernel(const uint8_t *__restrict__ buf, const uint64_t *__restrict__ poff,
                                                                       int64_t n_payload, int64_t n_frames, uint8_t *status,
                                                                       uint32_t *seq, uint8_t *chid, uint32_t *valid,
                                                                       unsigned long long *n_bad, uint8_t *imdt_spec, int skip)
{
    __shared__ __align__(16) uint32_t s_w[(IMTR_FRONT + IMTR_RUNS * IMTR_RUN_WORDS * 4 + 32) / 4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * IMTR_BATCH;
    const int n_here = (int)min((int64_t)IMTR_BATCH, n_frames - f0);
    const int64_t s_first = (int64_t)skip + f0 * 882;
    const int64_t i0 = s_first / 880;
    const int o0 = (int)(s_first - i0 * 880);              // frame 0 of the batch starts o0 bytes into run i0
    const int n_runs = (o0 + 882 * n_here + 879) / 880;      // runs that hold bytes of this batch (<= 34, all < n_payload)
    const unsigned long long pl0 = poff[min(i0 + lane, n_payload - 1)], pl1 = poff[min(i0 + 32 + lane, n_payload - 1)];
    uint32_t *runs = s_w + IMTR_FRONT / 4;
    // A warp copies runs wid, wid + 4, ...: lane l loads the ALIGNED source words l, l + 32, ... (221 words cover the run at any
    // alignment; no word lies entirely outside the payload, so nothing past the caller's buffer is ever touched); the upper
    // neighbour of a word comes from lane + 1 by shuffle, so a run costs 7 loads and 7 registers per lane and the loads of
    // IMTR_DEPTH runs are in flight while the oldest is shifted and stored (the first version loaded both words of every
    // pair and waited for one run at a time: 9 dependent memory round trips per warp, long-scoreboard stall 4.3 per issue).
    uint32_t R[IMTR_DEPTH][7], shv[IMTR_DEPTH];
    auto load_run = [&](int k, uint32_t (&dstw)[7], uint32_t &sh) {
        const unsigned long long a0 = __shfl_sync(0xffffffffu, pl0, k & 31), a1 = __shfl_sync(0xffffffffu, pl1, k & 31);
        const uint8_t *a = buf + (k < 32 ? a0 : a1);
        sh = (uint32_t)((uintptr_t)a & 3u);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a - sh);
        // an aligned run needs no word past its end: every word that is loaded holds at least one byte of the payload
        const int last = IMTR_RUN_WORDS - 1 + (sh ? 1 : 0);
#pragma unroll
        for (int u = 0; u < 7; ++u) dstw[u] = __ldg(w + min(lane + 32 * u, last));
    };
#pragma unroll
    for (int d = 0; d < IMTR_DEPTH - 1; ++d)
        if (wid + 4 * d < n_runs) load_run(wid + 4 * d, R[d], shv[d]);
    // (the ring index is a compile-time constant inside the IMTR_DEPTH-times unrolled body; the outer loop stays rolled: with
    // all nine iterations unrolled the kernel grew to 4096 instructions and stalled on instruction fetch, 1.8 per issue)
#pragma unroll 1
    for (int it0 = 0; wid + 4 * it0 < n_runs; it0 += IMTR_DEPTH) {
#pragma unroll
        for (int d = 0; d < IMTR_DEPTH; ++d) {
            const int k = wid + 4 * (it0 + d);
            if (k >= n_runs) break;
            if (k + 4 * (IMTR_DEPTH - 1) < n_runs) load_run(k + 4 * (IMTR_DEPTH - 1), R[(d + IMTR_DEPTH - 1) % IMTR_DEPTH], shv[(d + IMTR_DEPTH - 1) % IMTR_DEPTH]);
            uint32_t (&lo)[7] = R[d];
            const uint32_t sh8 = 8u * shv[d];
            uint32_t *dst = runs + IMTR_RUN_WORDS * k;
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                uint32_t hi = __shfl_down_sync(0xffffffffu, lo[u], 1);
                if (u < 6) {
                    const uint32_t first = __shfl_sync(0xffffffffu, lo[u + 1], 0);
                    hi = lane == 31 ? first : hi;
                }
                if (lane + 32 * u < IMTR_RUN_WORDS) dst[lane + 32 * u] = __funnelshift_r(lo[u], hi, sh8);
            }
        }
    }
    __syncthreads();
    const uint32_t fbase = (uint32_t)(IMTR_FRONT + o0);    // byte offset of frame 0 inside s_w
    // ---- CRCs of the 32 frames, bit-sliced exactly as in imtr_validate_kernel; only the loads differ: the span of frame q
    //      starts at byte fbase - 20 + 882 q, so odd and even frames have their own alignment (two aligned words + shift)
    __shared__ uint32_t s_part[3][16][32];              // the validating warp keeps its own part in registers
    const int vwarp = (int)(blockIdx.x & 3);            // rotate the validating warp over the SM's four schedulers
    uint32_t P[16], Q[16];
    {
        const uint32_t c = fbase - (uint32_t)(bitslice::SPAN - 876) + 28u * (uint32_t)lane; // this lane's piece of frame 0 (IMTR_FRONT >= 20)
        const uint32_t r = c & 3u, t = r + 2u;
        const uint32_t base_e = smem_u32(s_w) + (c - r), sh_e = 8u * r;
        const uint32_t base_o = smem_u32(s_w) + (c - r) + (t >= 4u ? 4u : 0u), sh_o = 8u * (t & 3u);
        const int clear_bit = lane == 0 ? 8 * (bitslice::SPAN - 876) : -1;
        const int jlo = 2 * wid, jhi = min(2 * wid + 2, 7);
#pragma unroll
        for (int i = 0; i < 16; ++i) P[i] = 0u;
#pragma unroll 1
        for (int j = jlo; j < jhi; ++j) {
            uint32_t T[32];
            const uint32_t ae = base_e + 4u * (uint32_t)j, ao = base_o + 4u * (uint32_t)j;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const uint32_t a = ((q & 1) ? ao : ae) + (uint32_t)((882 * q) & ~3);
                T[q] = __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (q & 1) ? sh_o : sh_e);
            }
            bitslice::transpose32(T);
            bitslice::lfsr_word(P, T, clear_bit - 32 * j);
        }
        if (lane == 0 && 32 * jhi <= 8 * (bitslice::SPAN - 876)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) P[i] = 0u;
        }
        if (wid == 0) bitslice::mul_xpow<160>(P, Q);
        else if (wid == 1) bitslice::mul_xpow<96>(P, Q);
        else if (wid == 2) bitslice::mul_xpow<32>(P, Q);
        else {
#pragma unroll
            for (int i = 0; i < 16; ++i) Q[i] = P[i];
        }
        if (wid != vwarp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s_part[(wid - vwarp - 1) & 3][i][lane] = Q[i];
        }
    }
    __syncthreads();
    if (wid != vwarp && imdt_spec) {
        // speculative output (see imtr_validate_kernel): frame f's 866 payload bytes land at f * 866.  Done by the three warps
        // that have no part in the last phase, WHILE the fourth joins the CRCs and validates (stores next to ALU work; with
        // the stores in front of the CRC phase the CTA kept its shared memory for ~1.5 us with one warp running)
        for (int q = (wid - vwarp - 1) & 3; q < n_here; q += IMTR_T / 32 - 1) {
            uint8_t *d = imdt_spec + (uint64_t)(f0 + q) * 866;
            const uint32_t pb = fbase + 882u * (uint32_t)q + 10u;        // payload start inside s_w (bytes)
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(s_w);
            const int head = (int)((4u - (uint32_t)((uintptr_t)d & 3u)) & 3u);
            if (lane < head) d[lane] = sb[pb + lane];
            const int nw = (866 - head) >> 2;
            const uint32_t s0 = pb + (uint32_t)head;
            const uint32_t sh = 8u * (s0 & 3u);
            const uint32_t *sw = s_w + (s0 >> 2);
            uint32_t *dw = reinterpret_cast<uint32_t *>(d + head);
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int k = lane + 32 * u;
                if (k < nw) dw[k] = __funnelshift_r(sw[k], sw[k + 1], sh);
            }
            const int t0 = head + 4 * nw;
            if (lane < 866 - t0) d[t0 + lane] = sb[pb + t0 + lane];
        }
    }
    if (wid != vwarp) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) P[i] = Q[i] ^ s_part[0][i][lane] ^ s_part[1][i][lane] ^ s_part[2][i][lane];
    {
        auto x = [](uint32_t v, int sd, int) { return __shfl_xor_sync(0xffffffffu, v, sd); };
        bitslice::join_level<1>(P, lane, x);
        bitslice::join_level<2>(P, lane, x);
        bitslice::join_level<4>(P, lane, x);
        bitslice::join_level<8>(P, lane, x);
        bitslice::join_level<16>(P, lane, x);
    }
    // ValidateImtrFrame, checks in the reference's order (ref aux_separator.h:558-590); lane = frame
    const uint8_t *fr = reinterpret_cast<const uint8_t *>(s_w) + fbase + 882u * (uint32_t)lane;
    int st = 0;
    if (!(fr[0] == 0x49 && fr[1] == 0x54 && fr[2] == 0xCE && fr[3] == 0x1F)) st = 1;                  // :559
    else if (!(fr[878] == 0x2E && fr[879] == 0xE9 && fr[880] == 0xC8 && fr[881] == 0xFD)) st = 2;     // :563
    else if (fr[9] != 0x22) st = 3;                                                                  // :572
    const uint32_t want = ((uint32_t)fr[876] << 8) | fr[877];
    if (st == 0 && (bitslice::unslice(P, lane) ^ bitslice::init_term(876)) != want) st = 4;
    const int64_t f = f0 + lane;
    if (f < n_frames) {
        status[f] = (uint8_t)st;
        seq[f] = ((uint32_t)fr[4] << 24) | ((uint32_t)fr[5] << 16) | ((uint32_t)fr[6] << 8) | fr[7];  // :568-569 BE u32 at 4
        chid[f] = fr[8];
        valid[f] = st == 0;
        if (st) atomicAdd(&n_bad[st - 1], 1ull);
    }
}

// seq of the valid frames in order (to evaluate the "previous accepted seq" rules)
__global__ void imtr_compact_seq_kernel(const uint32_t *valid, const uint32_t *rank, const uint32_t *seq, int64_t n,
                                        uint32_t *seq_c)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n && valid[f]) seq_c[rank[f]] = seq[f];
}
// restart rule: the IMDT file is (re)created when the previously accepted frame had seq 0
// (lastImtrSeq == 0, ref :513-528); gap rule :530-533
// prev_in: sequence number accepted before the first frame of this call (0 at the start of a stream); < 0 = unknown (a
// shard whose predecessor is not known yet): the rules of the first valid frame are left to the caller
__global__ void imtr_rules_kernel(const uint32_t *seq_c, const uint32_t *n_valid, unsigned long long *restart_last,
                                  unsigned long long *n_restarts, unsigned long long *n_gaps, long long prev_in, unsigned long long *seq_first_last)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)*n_valid) return;
    if (k == 0) seq_first_last[0] = seq_c[0];
    if (k + 1 == (int64_t)*n_valid) seq_first_last[1] = seq_c[k];
    if (k == 0 && prev_in < 0) return;
    const uint32_t prev = k ? seq_c[k - 1] : (uint32_t)prev_in;
    if (prev == 0u) {
        atomicMax(restart_last, (unsigned long long)k);
        atomicAdd(n_restarts, 1ull);
    }
    if (prev + 1u != seq_c[k]) atomicAdd(n_gaps, 1ull);
}
__global__ void __launch_bounds__(256) imtr_copy_kernel(const uint8_t *__restrict__ buf, const uint64_t *__restrict__ poff,
                                                        int64_t n_payload, int64_t n_frames, const uint32_t *valid,
                                                        const uint32_t *rank, const unsigned long long *restart_last,
                                                        const uint8_t *chid, uint8_t *imdt, uint64_t cap, int *first_chid,
                                                        const uint32_t *n_valid, int speculative, int skip)
{
    const int lane = threadIdx.x & 31;
    const uint64_t r0 = *restart_last;
    if (speculative && r0 == 0 && (int64_t)*n_valid == n_frames) { // imtr_validate_kernel already wrote every frame in place
        if (blockIdx.x == 0 && threadIdx.x == 0) *first_chid = chid[0];
        return;
    }
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n_frames; f += n_warps) {
    if (!valid[f]) continue;
    if (rank[f] < r0) continue;
    const uint64_t dst = (uint64_t)(rank[f] - r0) * 866;
    if (dst + 866 > cap) continue;
    if (rank[f] == r0 && lane == 0) *first_chid = chid[f];
    // the frames in front of the first rejected one are where the validating kernel put them (f * 866): one damaged
    // frame late in a downlink does not cost a second pass over everything before it
    if (speculative && r0 == 0 && (int64_t)rank[f] == f) continue;
    const FrameSegs S = frame_segs(buf, poff, n_payload, (int64_t)skip + f * 882);
    // 866 payload bytes from frame position 10 (IMTR_IMGDATA_OFF :72): aligned destination words, source words
    // re-aligned by funnel shift
    uint8_t *d = imdt + dst;
    const int head = (int)((4u - (uint32_t)((uintptr_t)d & 3u)) & 3u);
    if (lane < head) d[lane] = __ldg(seg_ptr(S, 10 + lane));
    const int nw = (866 - head) >> 2; // 215 or 216
    uint32_t *dw = reinterpret_cast<uint32_t *>(d + head);
    uint32_t v[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) {
        const int k = lane + 32 * u, q = 10 + head + 4 * k;
        v[u] = k < nw ? (seg_straddles(S, q) ? seg_word_bytes(S, q, 4) : seg_word(S, q)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 7; ++u) {
        const int k = lane + 32 * u;
        if (k < nw) dw[k] = v[u];
    }
    const int t0 = head + 4 * nw;
    if (lane < 866 - t0) d[t0 + lane] = __ldg(seg_ptr(S, 10 + t0 + lane));
    }
}

// =============================================================================================
// image frames: trailer-signature search, trailer gather, sub-image unpack
// =============================================================================================
// All occurrences of a 4-byte signature.  A thread takes one ALIGNED 16-byte chunk (one 128-bit load; the chunk grid starts
// at the buffer address rounded down to 16, positions in front of the buffer are never reported) plus the first word of
// the next chunk (from lane + 1 by shuffle; lane 31 loads it).  Filter: a position can only match where byte p is the
// first AND byte p + 1 the second signature byte -- (w ^ first) | ((w >> 8 | next << 24) ^ second) has a zero byte exactly
// there, found with the zero-byte trick (no false negatives), five instructions per word and one branch per chunk that is
// taken for ~1 chunk in 4000 on arbitrary data (the one-byte filter of round 1 sent almost every warp into the slow path:
// 85 instructions per 16 positions, issue-bound at 63 % of the slots; `first` is kept for the signature of the call).
__global__ void __launch_bounds__(256) find_sig4_kernel(const uint8_t *__restrict__ buf, int64_t n, uint32_t sig_le, uint8_t first,
                                                        unsigned long long *hits, uint32_t cap, uint32_t *n_hits)
{
    (void)first;
    const int lane = threadIdx.x & 31;
    const uint32_t mis = (uint32_t)((uintptr_t)buf & 15u);
    const uint8_t *A = buf - mis;                               // aligned origin; stream position p sits at A + mis + p
    const int64_t end = (int64_t)mis + n;                       // bytes [mis, end) of the aligned image are the stream
    const int64_t n_chunks = (end + 15) >> 4;
    const uint32_t p0pat = 0x01010101u * (sig_le & 0xFFu), p1pat = 0x01010101u * ((sig_le >> 8) & 0xFFu);
    auto scan16 = [&](const uint32_t (&w)[5], int64_t q0) {   // the 16 positions of one chunk (w[4] = first word of the next chunk)
        uint32_t acc = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t t = (w[i] ^ p0pat) | (__funnelshift_r(w[i], w[i + 1], 8) ^ p1pat);
            acc |= (t - 0x01010101u) & ~t;
        }
        if (acc & 0x80808080u) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int64_t p = q0 + 4 * i + b - (int64_t)mis;
                    if (__funnelshift_r(w[i], w[i + 1], 8 * b) == sig_le && p >= 0 && p + 4 <= n) {
                        uint32_t k = atomicAdd(n_hits, 1u);
                        if (k < cap) hits[k] = (unsigned long long)p;
                    }
                }
        }
    };
    // a warp takes FIND_M x 32 consecutive chunks per iteration (all loads issued before the first use: one 16-byte load per
    // thread in flight left the kernel latency-bound at 46 % of the DRAM rate); the loop bounds are warp-uniform
    constexpr int FIND_M = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * FIND_M;
    for (int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * FIND_M; c0 < n_chunks; c0 += stride) {
        if (((c0 + 32 * FIND_M) << 4) + 4 <= end) {          // everything this warp touches lies inside the stream image
            uint4 v[FIND_M];
#pragma unroll
            for (int m = 0; m < FIND_M; ++m) v[m] = __ldg(reinterpret_cast<const uint4 *>(A + ((c0 + 32 * m + lane) << 4)));
            uint32_t tail = 0u;
            if (lane == 31) tail = __ldg(reinterpret_cast<const uint32_t *>(A + ((c0 + 32 * FIND_M) << 4)));
#pragma unroll
            for (int m = 0; m < FIND_M; ++m) {
                uint32_t w[5] = {v[m].x, v[m].y, v[m].z, v[m].w, 0u};
                w[4] = __shfl_down_sync(0xffffffffu, v[m].x, 1);
                const uint32_t nxt = m + 1 < FIND_M ? __shfl_sync(0xffffffffu, v[m + 1 < FIND_M ? m + 1 : m].x, 0) : tail;
                if (lane == 31) w[4] = nxt;
                scan16(w, (c0 + 32 * m + lane) << 4);
            }
            continue;
        }
        for (int m = 0; m < FIND_M; ++m) {                    // the last chunks of the stream: guarded byte loads
            const int64_t q0 = (c0 + 32 * m + lane) << 4;
            uint32_t w[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                uint32_t x = 0;
                for (int b = 0; b < 4; ++b) {
                    const int64_t q = q0 + 4 * i + b;
                    x |= (uint32_t)(q < end ? A[q] : 0) << (8 * b);
                }
                w[i] = x;
            }
            scan16(w, q0);
        }
    }
}

__global__ void gather_trailers_kernel(const uint8_t *buf, int64_t n, const unsigned long long *hits, uint32_t n_hits,
                                       uint8_t *out)
{
    const uint32_t h = blockIdx.x;
    if (h >= n_hits) return;
    for (int b = threadIdx.x; b < 172; b += blockDim.x) {
        int64_t q = (int64_t)hits[h] + b;
        out[(size_t)h * 172 + b] = q < n ? buf[q] : 0;
    }
}

struct UnpackParams {
    const uint8_t *imdt;
    const int64_t *tab; // per frame: frame_off, tile_off[40]
    uint8_t *aux;
    uint16_t *pan, *mss;
    int tile_cols, tile_lines;
    int64_t n_frames;
};
// grid: (frame, item) with item 0..39 = sub-image r*8+c, 40 = aux block (ref aux_separator.h:335-393)
__global__ void __launch_bounds__(256) unpack_frames_kernel(const __grid_constant__ UnpackParams P)
{
    const int64_t fr = blockIdx.x;
    const int item = blockIdx.y;
    const int TC = P.tile_cols, TL = P.tile_lines;
    const int64_t W = 8 * (int64_t)TC;
    const int64_t *ent = P.tab + fr * 41;
    const int64_t frame_off = ent[0];
    if (item == 40) {
        if (!P.aux || blockIdx.z != 0) return;
        const int64_t nb = 192 * (int64_t)TL;
        uint8_t *dst = P.aux + fr * nb;
        if (frame_off < 0) {
            for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = 0;
        } else {
            const uint8_t *src = P.imdt + frame_off;
            if (((((uintptr_t)src) | ((uintptr_t)dst)) & 3) == 0) {
                for (int64_t i = threadIdx.x; i < nb / 4; i += blockDim.x)
                    reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
            } else {
                for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = src[i];
            }
        }
        return;
    }
    const int r = item >> 3, c = item & 7;
    uint16_t *dst;
    if (r < 4) {
        if (!P.pan) return;
        dst = P.pan + (fr * 4 * TL + (int64_t)r * TL) * W + (int64_t)c * TC;
    } else {
        if (!P.mss) return;
        dst = P.mss + fr * TL * W + (int64_t)c * TC;
    }
    const int64_t toff = ent[1 + item];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (frame_off < 0 || toff < 0) { // zero-filled gap frame (ref :302-311)
        if (blockIdx.z != 0) return;
        for (int y = wid; y < TL; y += nw)
            for (int x = lane; x < TC; x += 32) dst[(int64_t)y * W + x] = 0;
        return;
    }
    const uint8_t *src = P.imdt + toff;
    // blockIdx.z splits the tile's lines so that a handful of frames still fills the GPU
    const int y_lo = (int)(((int64_t)TL * blockIdx.z) / gridDim.z), y_hi = (int)(((int64_t)TL * (blockIdx.z + 1)) / gridDim.z);
    const bool a4 = ((((uintptr_t)src) & 3) == 0) && (TC % 2 == 0) && ((((uintptr_t)dst) & 3) == 0);
    // fast form: 16-byte output chunks (4 source words each; the stream is only 4-byte aligned), 4 chunks in flight per lane
    const bool a16 = a4 && (TC % 8 == 0) && ((((uintptr_t)dst) & 15) == 0) && ((W * 2) % 16 == 0);
    for (int y = y_lo + wid; y < y_hi; y += nw) {
        const uint8_t *sl = src + (int64_t)y * TC * 2;
        uint16_t *dl = dst + (int64_t)y * W;
        if (a16) {
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(sl);
            uint4 *dq = reinterpret_cast<uint4 *>(dl);
            const int nq = TC / 8;
            for (int q0 = 0; q0 < nq; q0 += 128) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + lane + 32 * u;
                    if (q < nq) {
                        v[u].x = __ldg(sw + 4 * q);     v[u].y = __ldg(sw + 4 * q + 1);
                        v[u].z = __ldg(sw + 4 * q + 2); v[u].w = __ldg(sw + 4 * q + 3);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + lane + 32 * u;
                    if (q < nq)
                        stg_na_v4(dq + q, make_uint4(bswap16x2(v[u].x), bswap16x2(v[u].y), bswap16x2(v[u].z), bswap16x2(v[u].w))); // :387-392
                }
            }
        } else if (a4) {
            for (int x = lane; x < TC / 2; x += 32)
                reinterpret_cast<uint32_t *>(dl)[x] = bswap16x2(reinterpret_cast<const uint32_t *>(sl)[x]); // :387-392
        } else {
            for (int x = lane; x < TC; x += 32) dl[x] = (uint16_t)(((uint32_t)sl[2 * x] << 8) | sl[2 * x + 1]);
        }
    }
}

__global__ void crc16_batch_kernel(const uint8_t *__restrict__ buf, const uint64_t *__restrict__ off, int64_t n,
                                   const __grid_constant__ CrcPlan crc, uint16_t *out)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const uint8_t *p = buf + off[i];
    const uint32_t c = warp_crc16(crc, [&](int k) { return (uint32_t)p[k]; });
    if ((threadIdx.x & 31) == 0) out[i] = (uint16_t)c;
}

} // namespace frames
} // namespace oip

using namespace oip;
using namespace oip::frames;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int oip_crc16_batch(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_off, int64_t n_items, int len,
                               uint16_t *d_crc)
{
    OIP_CHECK_CTX(ctx);
    if (!d_buf || !d_off || !d_crc) return fail(OIP_E_INVALID, "oip_crc16_batch: null pointer");
    if (len < 0 || len > (1 << 20)) return fail(OIP_E_INVALID, "oip_crc16_batch: len=%d", len);
    if (n_items <= 0) return OIP_OK;
    const CrcPlan plan = make_crc_plan(len);
    const int64_t blocks = (n_items * 32 + 255) / 256;
    crc16_batch_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_buf, d_off, n_items, plan, d_crc);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

static int aos_scan_impl(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, size_t own_bytes, size_t carry_in, uint64_t *d_payload_off,
                         size_t cap, int64_t counters[3], int64_t *carry_out);

extern "C" int oip_aos_scan(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, uint64_t *d_payload_off, size_t cap,
                            int64_t counters[3])
{
    return aos_scan_impl(ctx, d_buf, n_bytes, n_bytes, 0, d_payload_off, cap, counters, nullptr);
}

extern "C" int oip_aos_scan_shard(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, size_t own_bytes, size_t carry_in,
                                  uint64_t *d_payload_off, size_t cap, int64_t counters[3], int64_t *carry_out)
{
    if (own_bytes > n_bytes || carry_in > 1023 + 1) return fail(OIP_E_INVALID, "oip_aos_scan_shard: own_bytes=%zu of %zu, carry_in=%zu", own_bytes, n_bytes, carry_in);
    return aos_scan_impl(ctx, d_buf, n_bytes, own_bytes, carry_in, d_payload_off, cap, counters, carry_out);
}

static int aos_scan_impl(oip_ctx *ctx, const uint8_t *d_buf, size_t n_bytes, size_t own_bytes, size_t carry_in, uint64_t *d_payload_off,
                         size_t cap, int64_t counters[3], int64_t *carry_out)
{
    OIP_CHECK_CTX(ctx);
    if (counters) counters[0] = counters[1] = counters[2] = 0;
    if (carry_out) *carry_out = (int64_t)std::max<size_t>(carry_in, own_bytes) - (int64_t)own_bytes;
    if (!d_buf && n_bytes) return fail(OIP_E_INVALID, "oip_aos_scan: null buffer");
    if (n_bytes < 1024) return OIP_OK; // ref :623
    const bool shard = own_bytes != n_bytes || carry_in != 0;
    if (n_bytes / 4 > 0x7fffffffull) return fail(OIP_E_INVALID, "oip_aos_scan: buffer too large for one call");
    const int64_t n = (int64_t)n_bytes;
    const int64_t n_chunks = (n + CH - 1) / CH;
    static const CrcPlan plan = make_crc_plan(890);

    uint32_t cand_cap = (uint32_t)std::min<int64_t>(n / 4 + 1, n / 512 + 4096);
    bool fused = ctx->aos_fused != 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        // scratch layout
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
        const size_t o_hdr = take(64);                       // cursor(u32) | total(u32) | counters(3 x u64) | .. | acc_total @40 | phase @44 | shadowed @48
        const size_t o_info = take((size_t)n_chunks * sizeof(ChunkInfo));
        const size_t o_cnt = take((size_t)n_chunks * 4);
        const size_t o_base = take((size_t)n_chunks * 4);
        const size_t o_coff = take((size_t)cand_cap * 8);
        const size_t o_cst = take((size_t)cand_cap);
        const size_t o_ooff = take((size_t)cand_cap * 8);
        const size_t o_ost = take((size_t)cand_cap);
        const size_t o_rs = take((size_t)cand_cap);
        const size_t o_acc = take((size_t)cand_cap * 4);
        const size_t o_rank = take((size_t)cand_cap * 4);
        const size_t o_scan = take(scan_scratch_elems(std::max<int64_t>(n_chunks, cand_cap)) * 4);
        int rc = ensure_scratch(ctx, o);
        if (rc) return rc;
        uint8_t *S = (uint8_t *)ctx->d_scratch;
        uint32_t *d_cursor = (uint32_t *)(S + o_hdr);
        uint32_t *d_total = d_cursor + 1;
        unsigned long long *d_counters = (unsigned long long *)(S + o_hdr + 8);
        uint32_t *d_acc_total = (uint32_t *)(S + o_hdr + 40);
        uint32_t *d_phase = (uint32_t *)(S + o_hdr + 44);
        uint32_t *d_shadowed = (uint32_t *)(S + o_hdr + 48);
        uint32_t *d_irregular = (uint32_t *)(S + o_hdr + 52);
        uint32_t *d_m_own = (uint32_t *)(S + o_hdr + 56);
        unsigned long long *d_last_end = (unsigned long long *)(S + o_hdr + 32);
        ChunkInfo *d_info = (ChunkInfo *)(S + o_info);
        OIP_CUDA(cudaMemsetAsync(S + o_hdr, 0, 64, ctx->stream));

        if (fused) {
            // one pass: cadence phase, then sync + rules + CRC of 32 slots per warp (CH == GRP: the same group table)
            OIP_CUDA(cudaMemsetAsync(d_phase, 0xFF, 4, ctx->stream));
            aos_phase_kernel<<<64, 256, 0, ctx->stream>>>(d_buf, n, d_phase);
            OIP_CUDA(cudaGetLastError());
            aos_fused_kernel<<<(unsigned)((n_chunks + AOS_F_WARPS - 1) / AOS_F_WARPS), AOS_F_WARPS * 32, 0, ctx->stream>>>(
                d_buf, n, d_phase, n_chunks, d_cursor, cand_cap, d_info, (uint64_t *)(S + o_coff), (int8_t *)(S + o_cst), d_irregular);
            OIP_CUDA(cudaGetLastError());
            ctx->launches += 2;
        } else {
            aos_scan_kernel<<<(unsigned)n_chunks, AOS_T, 0, ctx->stream>>>(d_buf, n, d_cursor, cand_cap, d_info,
                                                                          (uint64_t *)(S + o_coff), (int8_t *)(S + o_cst));
            OIP_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        rc = ensure_pinned(ctx, 64);
        if (rc) return rc;
        // everything below is sized by the table capacity and reads the candidate count on the device: one host round
        // trip per call (the count only decides whether the pathological retry is needed)
        // file order: scan the per-chunk counts (ChunkInfo.count is strided -> copy out first)
        OIP_CUDA(cudaMemcpy2DAsync(S + o_cnt, 4, (uint8_t *)d_info + 4, sizeof(ChunkInfo), 4, (size_t)n_chunks,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        rc = exclusive_scan_u32(ctx, (uint32_t *)(S + o_cnt), (uint32_t *)(S + o_base), n_chunks, (uint32_t *)(S + o_scan),
                                d_total);
        if (rc) return rc;
        aos_clamp_kernel<<<1, 1, 0, ctx->stream>>>(d_total, cand_cap);
        OIP_CUDA(cudaGetLastError());
        aos_order_kernel<<<(unsigned)((n_chunks * 32 + 255) / 256), 256, 0, ctx->stream>>>(
            d_info, (uint32_t *)(S + o_base), n_chunks, cand_cap, (uint64_t *)(S + o_coff), (int8_t *)(S + o_cst),
            (uint64_t *)(S + o_ooff), (int8_t *)(S + o_ost));
        OIP_CUDA(cudaGetLastError());
        // (after the fused kernel only the irregular candidates still wait for their CRC: false sync words inside rejected
        // frames, groups without cadence)
        aos_crc_kernel<<<(unsigned)(((size_t)cand_cap + 32 * AOS_CRC_WARPS - 1) / (32 * AOS_CRC_WARPS)), AOS_CRC_WARPS * 32, 0, ctx->stream>>>(
            d_buf, plan, (uint64_t *)(S + o_ooff), (int8_t *)(S + o_ost), d_total, fused ? d_irregular : nullptr);
        OIP_CUDA(cudaGetLastError());
        const unsigned gb = (unsigned)(((size_t)cand_cap + 255) / 256);
        OIP_CUDA(cudaMemsetAsync(S + o_acc, 0, (size_t)cand_cap * 4, ctx->stream)); // entries past the count stay 0 for the scan
        if (shard) { // from here on the table holds the shard's own candidates only (the total stays in d_cursor for the overflow test)
            aos_shard_trim_kernel<<<gb, 256, 0, ctx->stream>>>((uint64_t *)(S + o_ooff), (int8_t *)(S + o_ost), d_total, (uint64_t)own_bytes,
                                                               (uint64_t)carry_in, d_m_own);
            OIP_CUDA(cudaGetLastError());
            OIP_CUDA(cudaMemcpyAsync(d_total, d_m_own, 4, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->launches++;
        }
        aos_runstart_kernel<<<gb, 256, 0, ctx->stream>>>((uint64_t *)(S + o_ooff), (int8_t *)(S + o_ost), d_total, S + o_rs);
        OIP_CUDA(cudaGetLastError());
        aos_walk_kernel<<<gb, 256, 0, ctx->stream>>>((uint64_t *)(S + o_ooff), (int8_t *)(S + o_ost), S + o_rs, d_total,
                                                     (uint32_t *)(S + o_acc), d_counters, d_shadowed);
        OIP_CUDA(cudaGetLastError());
        ctx->launches += 5;
        rc = exclusive_scan_u32(ctx, (uint32_t *)(S + o_acc), (uint32_t *)(S + o_rank), cand_cap, (uint32_t *)(S + o_scan),
                                d_acc_total);
        if (rc) return rc;
        if (d_payload_off || carry_out) {
            aos_emit_kernel<<<gb, 256, 0, ctx->stream>>>((uint64_t *)(S + o_ooff), (uint32_t *)(S + o_acc),
                                                         (uint32_t *)(S + o_rank), d_total, d_payload_off, (uint64_t)cap, d_last_end);
            OIP_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        // header: cursor(u32) | total(u32) | counters(3 x u64) | .. | shadowed-valid flag @48
        uint8_t *hb = (uint8_t *)ctx->h_pinned;
        OIP_CUDA(cudaMemcpyAsync(hb, S + o_hdr, 64, cudaMemcpyDeviceToHost, ctx->stream));
        OIP_CUDA(cudaStreamSynchronize(ctx->stream));
        const uint32_t m = *(const uint32_t *)hb;
        if (m > cand_cap) { // pathological input (sync pattern everywhere): retry with the exact size
            cand_cap = m;
            continue;
        }
        if (fused && *(const uint32_t *)(hb + 48)) { // a valid frame shadowed by an overlapping one: exhaustive search
            fused = false;
            continue;
        }
        const unsigned long long *hc = (const unsigned long long *)(hb + 8);
        if (counters) { counters[0] = (int64_t)hc[0]; counters[1] = (int64_t)hc[1]; counters[2] = (int64_t)hc[2]; }
        if (carry_out) { // where the next shard's scan starts, relative to its first byte
            const unsigned long long le = *(const unsigned long long *)(hb + 32);
            *carry_out = (int64_t)std::max<unsigned long long>(std::max<size_t>(carry_in, own_bytes), le) - (int64_t)own_bytes;
        }
        if (d_payload_off && hc[0] > cap) return fail(OIP_E_INVALID, "oip_aos_scan: %llu valid frames exceed capacity %zu", hc[0], cap);
        return OIP_OK;
    }
    return fail(OIP_E_NOMEM, "oip_aos_scan: candidate table overflow");
}

static int imtr_deframe_impl(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off, int64_t n_payload, int skip,
                             int64_t nf, long long prev_seq, uint8_t *d_imdt, size_t cap, int64_t stats[9], int64_t *imdt_bytes,
                             int64_t seq_info[3]);

extern "C" int oip_imtr_deframe(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off, int64_t n_payload,
                                uint8_t *d_imdt, size_t cap, int64_t stats[9], int64_t *imdt_bytes)
{
    if (n_payload < 0) return fail(OIP_E_INVALID, "oip_imtr_deframe: n_payload < 0");
    return imtr_deframe_impl(ctx, d_buf, d_payload_off, n_payload, 0, n_payload * 880 / 882 /* frames cut by the cadence, ref :487-510 */, 0,
                             d_imdt, cap, stats, imdt_bytes, nullptr);
}

extern "C" int oip_imtr_deframe_shard(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off, int64_t n_payload,
                                      int skip_bytes, int64_t n_frames, int64_t prev_seq, uint8_t *d_imdt, size_t cap, int64_t stats[9],
                                      int64_t *imdt_bytes, int64_t seq_info[3])
{
    if (n_payload < 0 || n_frames < 0 || skip_bytes < 0 || skip_bytes > 881 || (int64_t)skip_bytes + n_frames * 882 > n_payload * 880)
        return fail(OIP_E_INVALID, "oip_imtr_deframe_shard: %lld frames from byte %d do not fit %lld payloads", (long long)n_frames, skip_bytes,
                    (long long)n_payload);
    return imtr_deframe_impl(ctx, d_buf, d_payload_off, n_payload, skip_bytes, n_frames, prev_seq, d_imdt, cap, stats, imdt_bytes, seq_info);
}

static int imtr_deframe_impl(oip_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_payload_off, int64_t n_payload, int skip,
                             int64_t nf, long long prev_seq, uint8_t *d_imdt, size_t cap, int64_t stats[9], int64_t *imdt_bytes,
                             int64_t seq_info[3])
{
    OIP_CHECK_CTX(ctx);
    if (stats) { for (int i = 0; i < 9; ++i) stats[i] = 0; stats[7] = -1; }
    if (imdt_bytes) *imdt_bytes = 0;
    if (seq_info) { seq_info[0] = seq_info[1] = -1; seq_info[2] = -1; }
    if (nf == 0) return OIP_OK;
    if (!d_buf || !d_payload_off || !d_imdt) return fail(OIP_E_INVALID, "oip_imtr_deframe: null pointer");
    if (nf > 0x7fffffff) return fail(OIP_E_INVALID, "oip_imtr_deframe: too many frames for one call");
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_hdr = take(96); // restart_last | n_restarts | n_gaps (u64 each) | total(u32) | first_chid(i32) | n_bad[4] (u64) | seq first, last (u64)
    const size_t o_status = take((size_t)nf);
    const size_t o_chid = take((size_t)nf);
    const size_t o_seq = take((size_t)nf * 4);
    const size_t o_valid = take((size_t)nf * 4);
    const size_t o_rank = take((size_t)nf * 4);
    const size_t o_seqc = take((size_t)nf * 4);
    const size_t o_scan = take(scan_scratch_elems(nf) * 4);
    int rc = ensure_scratch(ctx, o);
    if (rc) return rc;
    rc = ensure_pinned(ctx, 96);
    if (rc) return rc;
    uint8_t *S = (uint8_t *)ctx->d_scratch;
    unsigned long long *d_hdr = (unsigned long long *)(S + o_hdr);
    uint32_t *d_total = (uint32_t *)(S + o_hdr + 24);
    int *d_first_chid = (int *)(S + o_hdr + 28);
    unsigned long long *d_bad = (unsigned long long *)(S + o_hdr + 32);
    OIP_CUDA(cudaMemsetAsync(S + o_hdr, 0, 96, ctx->stream));
    OIP_CUDA(cudaMemsetAsync(d_first_chid, 0xFF, 4, ctx->stream));
    // one stream-ordered chain, one host round trip at the end
    const bool speculative = (uint64_t)nf * 866 <= (uint64_t)cap; // room for every cut frame: validate writes them in place
    if (!ctx->imtr_attr_set) { // six 36 KB CTAs per SM need the large shared-memory carve-out
        OIP_CUDA(cudaFuncSetAttribute(imtr_validate_runs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        ctx->imtr_attr_set = true;
    }
    (ctx->imtr_runs ? imtr_validate_runs_kernel : imtr_validate_kernel)<<<(unsigned)((nf + IMTR_BATCH - 1) / IMTR_BATCH), IMTR_T, 0, ctx->stream>>>(
        d_buf, d_payload_off, n_payload, nf, S + o_status, (uint32_t *)(S + o_seq), S + o_chid, (uint32_t *)(S + o_valid), d_bad,
        speculative ? d_imdt : nullptr, skip);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    rc = exclusive_scan_u32(ctx, (uint32_t *)(S + o_valid), (uint32_t *)(S + o_rank), nf, (uint32_t *)(S + o_scan), d_total);
    if (rc) return rc;
    const unsigned gb = (unsigned)((nf + 255) / 256);
    imtr_compact_seq_kernel<<<gb, 256, 0, ctx->stream>>>((uint32_t *)(S + o_valid), (uint32_t *)(S + o_rank),
                                                         (uint32_t *)(S + o_seq), nf, (uint32_t *)(S + o_seqc));
    OIP_CUDA(cudaGetLastError());
    imtr_rules_kernel<<<gb, 256, 0, ctx->stream>>>((uint32_t *)(S + o_seqc), d_total, d_hdr, d_hdr + 1, d_hdr + 2, prev_seq, d_hdr + 8);
    OIP_CUDA(cudaGetLastError());
    imtr_copy_kernel<<<(unsigned)std::min<int64_t>((nf * 32 + 255) / 256, (int64_t)ctx->sm_count * 32), 256, 0, ctx->stream>>>(
        d_buf, d_payload_off, n_payload, nf, (uint32_t *)(S + o_valid), (uint32_t *)(S + o_rank), d_hdr, S + o_chid, d_imdt,
        (uint64_t)cap, d_first_chid, d_total, speculative ? 1 : 0, skip);
    OIP_CUDA(cudaGetLastError());
    ctx->launches += 3;
    uint8_t *hp = (uint8_t *)ctx->h_pinned;
    OIP_CUDA(cudaMemcpyAsync(hp, S + o_hdr, 96, cudaMemcpyDeviceToHost, ctx->stream));
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long *hh = (const unsigned long long *)hp;
    const uint32_t n_valid = *(const uint32_t *)(hp + 24);
    const int64_t bad[5] = {0, (int64_t)hh[4], (int64_t)hh[5], (int64_t)hh[6], (int64_t)hh[7]};
    const int64_t out_frames = n_valid ? (int64_t)n_valid - (int64_t)hh[0] : 0;
    if ((uint64_t)out_frames * 866 > cap) return fail(OIP_E_INVALID, "oip_imtr_deframe: output capacity %zu too small", cap);
    if (stats) {
        stats[0] = nf; stats[1] = n_valid; stats[2] = bad[1]; stats[3] = bad[2]; stats[4] = bad[3]; stats[5] = bad[4];
        stats[6] = (int64_t)hh[2]; stats[7] = n_valid ? *(int *)(hp + 28) : -1; stats[8] = (int64_t)hh[1];
    }
    if (seq_info && n_valid) {
        seq_info[0] = (int64_t)hh[8]; seq_info[1] = (int64_t)hh[9];
        seq_info[2] = hh[1] ? (int64_t)hh[0] : -1; // index (among this call's valid frames) of the last restart, -1: none
    }
    if (imdt_bytes) *imdt_bytes = out_frames * 866;
    return OIP_OK;
}

// device part of the frame index: every occurrence of EB 90 E1 4D, in ascending order, with the 172 bytes that start there
static int frames_hits(oip_ctx *ctx, const uint8_t *d_imdt, int64_t n, std::vector<unsigned long long> &hits_sorted, std::vector<uint8_t> &trailers_sorted)
{
    uint32_t hit_cap = (uint32_t)std::min<int64_t>(n / 4 + 1, std::max<int64_t>(4096, n / 65536));
    std::vector<unsigned long long> hits;
    std::vector<uint8_t> trailers;
    for (int attempt = 0; attempt < 2; ++attempt) {
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
        const size_t o_n = take(16), o_hits = take((size_t)hit_cap * 8), o_tr = take((size_t)hit_cap * 172);
        int rc = ensure_scratch(ctx, o);
        if (rc) return rc;
        uint8_t *S = (uint8_t *)ctx->d_scratch;
        OIP_CUDA(cudaMemsetAsync(S + o_n, 0, 16, ctx->stream));
        const int blocks = (int)std::min<int64_t>((n / 64 + 1 + 255) / 256 + 1, (int64_t)ctx->sm_count * 16);
        find_sig4_kernel<<<blocks, 256, 0, ctx->stream>>>(d_imdt, n, 0x4DE190EBu, 0xEB, (unsigned long long *)(S + o_hits),
                                                         hit_cap, (uint32_t *)(S + o_n));
        OIP_CUDA(cudaGetLastError());
        ctx->launches++;
        uint32_t nh = 0;
        OIP_CUDA(cudaMemcpyAsync(&nh, S + o_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
        OIP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (nh > hit_cap) { hit_cap = nh; continue; }
        hits.resize(nh);
        trailers.resize((size_t)nh * 172);
        if (nh) {
            gather_trailers_kernel<<<nh, 64, 0, ctx->stream>>>(d_imdt, n, (unsigned long long *)(S + o_hits), nh, S + o_tr);
            OIP_CUDA(cudaGetLastError());
            ctx->launches++;
            OIP_CUDA(cudaMemcpyAsync(hits.data(), S + o_hits, (size_t)nh * 8, cudaMemcpyDeviceToHost, ctx->stream));
            OIP_CUDA(cudaMemcpyAsync(trailers.data(), S + o_tr, (size_t)nh * 172, cudaMemcpyDeviceToHost, ctx->stream));
            OIP_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        break;
    }
    // order by offset (atomics give arbitrary order)
    std::vector<uint32_t> order(hits.size());
    for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hits[a] < hits[b]; });
    hits_sorted.resize(hits.size());
    trailers_sorted.resize(trailers.size());
    for (size_t i = 0; i < order.size(); ++i) {
        hits_sorted[i] = hits[order[i]];
        memcpy(trailers_sorted.data() + i * 172, trailers.data() + (size_t)order[i] * 172, 172);
    }
    return OIP_OK;
}

extern "C" int oip_image_frames_hits(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes, uint64_t *hits, uint8_t *trailers, int64_t cap,
                                     int64_t *n_hits)
{
    OIP_CHECK_CTX(ctx);
    if (!n_hits) return fail(OIP_E_INVALID, "oip_image_frames_hits: n_hits is null");
    *n_hits = 0;
    if (!d_imdt && n_bytes) return fail(OIP_E_INVALID, "oip_image_frames_hits: null buffer");
    if (n_bytes < 4) return OIP_OK;
    std::vector<unsigned long long> h;
    std::vector<uint8_t> t;
    int rc = frames_hits(ctx, d_imdt, (int64_t)n_bytes, h, t);
    if (rc) return rc;
    *n_hits = (int64_t)h.size();
    if ((int64_t)h.size() > cap) return fail(OIP_E_INVALID, "oip_image_frames_hits: %zu signatures exceed capacity %lld", h.size(), (long long)cap);
    if (h.size() && (!hits || !trailers)) return fail(OIP_E_INVALID, "oip_image_frames_hits: null output");
    for (size_t i = 0; i < h.size(); ++i) hits[i] = (uint64_t)h[i];
    if (t.size()) memcpy(trailers, t.data(), t.size());
    return OIP_OK;
}

// host part of the frame index (no device, no context): NextImageDataFrame chain + gap rules over the signature table
extern "C" int oip_image_frames_chain(const uint64_t *hits, const uint8_t *trailer_bytes, int64_t n_hits, size_t n_bytes, const oip_frame_geom *geom,
                                      oip_frame_entry *entries, int64_t cap, int64_t stats[4])
{
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (!geom || geom->tile_cols < 1 || geom->tile_lines < 1) return fail(OIP_E_INVALID, "oip_image_frames_chain: bad geometry");
    if (n_hits < 0 || (n_hits && (!hits || !trailer_bytes))) return fail(OIP_E_INVALID, "oip_image_frames_chain: bad signature table");
    const int64_t n = (int64_t)n_bytes;
    const int64_t aux_all = 192 * (int64_t)geom->tile_lines;            // IMGSIG_AUX_ALLBYTES
    const int64_t tile_bytes = (int64_t)geom->tile_lines * geom->tile_cols * 2;
    if (n <= aux_all + 172) return OIP_OK;                               // ref :630
    // ---- host: NextImageDataFrame chain + gap rules (ref :627-656, :287-320); only trailers are needed
    auto be32 = [](const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
    int64_t p = 0, remain = n, emitted = 0, found = 0, incomplete = 0;
    int last_seq = 0;
    size_t hi = 0;
    for (;;) {
        if (remain <= aux_all + 172) break;                              // :630
        while (hi < (size_t)n_hits && (int64_t)hits[hi] < p) ++hi;        // memmem from p
        if (hi >= (size_t)n_hits) break;
        const int64_t sp = (int64_t)hits[hi];
        if (sp + 172 > n) break; // the reference would parse past the mapping
        const uint8_t *t = trailer_bytes + hi * 172;
        const int64_t frame_end = sp + 172;                              // :634
        const int z_ratio = t[4] & 0x3F;                                 // :639
        const int seq = (int)(((uint32_t)t[6] << 8) | t[7]);            // :642-643
        const uint32_t image_dwords = be32(t + 8);                       // :645-646
        const int data_bytes = (int)(uint32_t)((uint64_t)image_dwords * 4u + (uint64_t)aux_all); // :653
        if (sp - p < (int64_t)data_bytes) {                              // :654, :289-299
            incomplete++;
            remain -= frame_end - p;
            p = frame_end;
            continue;
        }
        if (data_bytes < 0) return fail(OIP_E_RANGE, "image frame #%d: trailer length field out of range", seq);
        const int64_t frame = sp - data_bytes;                           // :655
        found++;
        if (z_ratio != 0)
            return fail(OIP_E_UNSUPPORTED, "image frame #%d is JPEG-2000 compressed (z_ratio 0x%02X): not supported", seq, z_ratio);
        if (seq > last_seq + 1) {                                        // :302-311
            for (int i = 0; i < seq - last_seq - 1; ++i) {
                if (emitted < cap && entries) {
                    oip_frame_entry &e = entries[emitted];
                    e.frame_off = -1;
                    for (int k = 0; k < 40; ++k) e.tile_off[k] = -1;
                    e.seq = last_seq + 1 + i;
                    e.z_ratio = 0;
                }
                emitted++;
            }
        }
        if (emitted < cap && entries) {
            oip_frame_entry &e = entries[emitted];
            e.frame_off = frame;
            e.seq = seq;
            e.z_ratio = z_ratio;
            int64_t q = frame + aux_all;
            for (int k = 0; k < 40; ++k) {                               // :347-356
                if (q + tile_bytes > n) return fail(OIP_E_RANGE, "image frame #%d: sub-image %d leaves the buffer", seq, k);
                e.tile_off[k] = q;
                q += (int64_t)be32(t + 12 + 4 * k) * 4;
            }
        }
        emitted++;
        remain -= frame_end - p;                                         // :315-317
        p = frame_end;
        last_seq = seq;
    }
    if (stats) { stats[0] = found; stats[1] = emitted; stats[2] = incomplete; stats[3] = last_seq; }
    if (entries && emitted > cap) return fail(OIP_E_INVALID, "oip_image_frames_index: %lld frames exceed capacity %lld", (long long)emitted, (long long)cap);
    return OIP_OK;
}

extern "C" int oip_image_frames_index(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes, const oip_frame_geom *geom,
                                      oip_frame_entry *entries, int64_t cap, int64_t stats[4])
{
    OIP_CHECK_CTX(ctx);
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (!geom || geom->tile_cols < 1 || geom->tile_lines < 1) return fail(OIP_E_INVALID, "oip_image_frames_index: bad geometry");
    if (!d_imdt && n_bytes) return fail(OIP_E_INVALID, "oip_image_frames_index: null buffer");
    if ((int64_t)n_bytes <= 192 * (int64_t)geom->tile_lines + 172) return OIP_OK;                    // ref :630
    std::vector<unsigned long long> h;
    std::vector<uint8_t> t;
    int rc = frames_hits(ctx, d_imdt, (int64_t)n_bytes, h, t);
    if (rc) return rc;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "hit table layout");
    return oip_image_frames_chain(reinterpret_cast<const uint64_t *>(h.data()), t.data(), (int64_t)h.size(), n_bytes, geom, entries, cap, stats);
}

extern "C" int oip_unpack_frames(oip_ctx *ctx, const uint8_t *d_imdt, size_t n_bytes, const oip_frame_geom *geom,
                                 const oip_frame_entry *entries, int64_t n_frames, uint8_t *d_aux, uint16_t *d_pan,
                                 uint16_t *d_mss)
{
    OIP_CHECK_CTX(ctx);
    if (!geom || geom->tile_cols < 1 || geom->tile_lines < 1) return fail(OIP_E_INVALID, "oip_unpack_frames: bad geometry");
    if (n_frames < 0 || (n_frames && !entries)) return fail(OIP_E_INVALID, "oip_unpack_frames: bad entries");
    if (n_frames == 0) return OIP_OK;
    if (n_frames > 0x7fffffff) return fail(OIP_E_INVALID, "oip_unpack_frames: too many frames");
    const int64_t tile_bytes = (int64_t)geom->tile_lines * geom->tile_cols * 2;
    const int64_t aux_all = 192 * (int64_t)geom->tile_lines;
    const size_t tab_bytes = (size_t)n_frames * 41 * 8;
    int rc = ensure_pinned(ctx, tab_bytes);
    if (rc) return rc;
    rc = ensure_scratch(ctx, tab_bytes);
    if (rc) return rc;
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t *tab = (int64_t *)ctx->h_pinned;
    for (int64_t f = 0; f < n_frames; ++f) {
        const oip_frame_entry &e = entries[f];
        if (e.frame_off >= 0) {
            if (e.z_ratio != 0) return fail(OIP_E_UNSUPPORTED, "frame %lld is compressed", (long long)f);
            if (e.frame_off + aux_all > (int64_t)n_bytes) return fail(OIP_E_RANGE, "frame %lld aux block leaves the buffer", (long long)f);
            for (int k = 0; k < 40; ++k)
                if (e.tile_off[k] < 0 || e.tile_off[k] + tile_bytes > (int64_t)n_bytes)
                    return fail(OIP_E_RANGE, "frame %lld sub-image %d leaves the buffer", (long long)f, k);
        }
        tab[f * 41] = e.frame_off;
        for (int k = 0; k < 40; ++k) tab[f * 41 + 1 + k] = e.tile_off[k];
    }
    OIP_CUDA(cudaMemcpyAsync(ctx->d_scratch, tab, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
    UnpackParams P{};
    P.imdt = d_imdt; P.tab = (const int64_t *)ctx->d_scratch; P.aux = d_aux; P.pan = d_pan; P.mss = d_mss;
    P.tile_cols = geom->tile_cols; P.tile_lines = geom->tile_lines; P.n_frames = n_frames;
    const unsigned zsplit = (unsigned)std::max(1, std::min(geom->tile_lines / 8, (int)(4096 / std::max<int64_t>(1, n_frames * 41) + 1)));
    unpack_frames_kernel<<<dim3((unsigned)n_frames, 41, zsplit), 256, 0, ctx->stream>>>(P);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}
