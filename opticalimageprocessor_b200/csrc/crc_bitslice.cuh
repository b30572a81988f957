// crc_bitslice.cuh -- CRC-16/CCITT-FALSE of 32 frames at once in bit-sliced form (ref CRC.h:806-834 is bit-serial;
// parameters ref CRC.h:1522-1526: poly 0x1021, init 0xFFFF, MSB first, no reflection, xorout 0).
//
// Register i of the sliced state holds bit i of the CRC register of 32 different frames (bit f = frame f), so one
// step of the reference's bit-serial recurrence costs three XORs for 32 frames: 0.1 instructions per message
// byte and frame instead of ~8 for the byte-wise form.  The price is a 32x32 bit transpose per 32 loaded words.
//
//   * the 896-byte span that ENDS with the last message byte is cut into 32 pieces of 28 bytes, one per lane; the
//     bytes in front of the message (HEAD = 896 - len) belong to lane 0, which clears its state after them (a zero
//     remainder is not changed by whatever was shifted through before it is cleared)
//   * every lane runs its 7 words through transpose32 + 32 recurrence steps each, from a zero remainder
//   * the 32 partial remainders are joined by a butterfly: at distance s the earlier partner's value is multiplied by
//     x^(224*s) mod P -- a fixed 16x16 GF(2) matrix, i.e. a compile-time XOR network on the sliced registers
//   * the 0xFFFF initial value contributes 0xFFFF * x^(8*len) mod P, a constant per frame format
//
// Host-compilable (plain C++17) so the algorithm is unit-tested on the CPU (tests/test_crc_bitslice_cpu.py); the
// warp exchange of the butterfly is the only device-specific part.
#pragma once
#include <cstdint>
#include <utility>

#if defined(__CUDACC__)
#define OIP_BS_HD __host__ __device__ __forceinline__
#else
#define OIP_BS_HD inline
#endif

namespace oip {
namespace bitslice {

constexpr int PIECE = 28;          // message bytes per lane
constexpr int SPAN = 32 * PIECE;   // 896 bytes covered by a warp

constexpr uint16_t gf_mulx(uint16_t r) { return (uint16_t)((r & 0x8000) ? ((r << 1) ^ 0x1021) : (r << 1)); }
constexpr uint16_t gf_xpow(int nbits) // x^nbits mod P
{
    uint16_t r = 1;
    for (int i = 0; i < nbits; ++i) r = gf_mulx(r);
    return r;
}
constexpr uint16_t gf_mul(uint16_t a, uint16_t b)
{
    uint16_t r = 0;
    for (int i = 15; i >= 0; --i) {
        r = gf_mulx(r);
        if ((b >> i) & 1) r ^= a;
    }
    return r;
}
// contribution of the 0xFFFF initial remainder to a message of len bytes
constexpr uint16_t init_term(int len) { return gf_mul(0xFFFF, gf_xpow(8 * len)); }

// column j of the matrix of "multiply by x^NBITS mod P": (x^j * x^NBITS) mod P
template <int NBITS> constexpr uint16_t mul_col(int j)
{
    uint16_t c = gf_xpow(NBITS);
    for (int i = 0; i < j; ++i) c = gf_mulx(c);
    return c;
}

// 32x32 bit-matrix transpose in place: afterwards bit f of a[k] = bit k of the original a[f]
OIP_BS_HD void transpose32(uint32_t (&a)[32])
{
#if defined(__CUDA_ARCH__)
    // distance 16 and 8 move whole halfwords / bytes: one PRMT per output word
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t lo = a[i], hi = a[i + 16];
        a[i] = __byte_perm(lo, hi, 0x5410);
        a[i + 16] = __byte_perm(lo, hi, 0x7632);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i & 8) continue;
        const uint32_t lo = a[i], hi = a[i + 8];
        a[i] = __byte_perm(lo, hi, 0x6240);
        a[i + 8] = __byte_perm(lo, hi, 0x7351);
    }
#else
    for (int s = 16; s >= 8; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000FFFFu : 0x00FF00FFu;
        for (int i = 0; i < 32; ++i) {
            if (i & s) continue;
            const uint32_t t = ((a[i] >> s) ^ a[i + s]) & m;
            a[i + s] ^= t;
            a[i] ^= t << s;
        }
    }
#endif
    // (round 2 measured the two-bit-select form of these stages -- (x & ~m) | (y & m), four instructions per pair instead of
    // the five of the xor swap: aos_fused_kernel 260 -> 279 us, imtr_validate_runs_kernel 423 -> 443 us; not kept)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 4; s >= 1; s >>= 1) {
        const uint32_t m = s == 4 ? 0x0F0F0F0Fu : (s == 2 ? 0x33333333u : 0x55555555u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 32; ++i) {
            if (i & s) continue;
            const uint32_t t = ((a[i] >> s) ^ a[i + s]) & m;
            a[i + s] ^= t;
            a[i] ^= t << s;
        }
    }
}

// 32 steps of the bit-serial recurrence on the sliced state for one transposed message word (loaded little-endian:
// message order is byte 0 bit 7..0, byte 1 bit 7..0, ...).  The shift of the register is a renaming: logical bit i
// lives in P[(i - t) & 15] before step t, so 16 steps return to the identity.  When t reaches clear_at (0 or 16, or
// anything else for "never") the state is zeroed first -- lane 0 drops the bytes in front of the message.
OIP_BS_HD void lfsr_word(uint32_t (&P)[16], const uint32_t (&T)[32], int clear_at)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < 32; ++t) {
        if (t == 0 || t == 16) {
            if (clear_at == t) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 16; ++i) P[i] = 0u;
            }
        }
        const int k = 8 * (t >> 3) + 7 - (t & 7);
        const int a = (15 - t) & 15, b12 = (11 - t) & 15, b5 = (4 - t) & 15;
        const uint32_t fb = P[a] ^ T[k]; // bit shifted out ^ message bit
        P[a] = fb;                       // becomes logical bit 0
        P[b12] ^= fb;                    // logical bit 11 -> 12, tap x^12
        P[b5] ^= fb;                     // logical bit 4 -> 5, tap x^5
    }
}

// out = in * x^NBITS mod P on sliced registers: a compile-time XOR network
template <int NBITS, int I, int J> struct MulBit {
    static constexpr bool value = ((mul_col<NBITS>(J) >> I) & 1) != 0;
};
template <int NBITS, int I, int... J> OIP_BS_HD uint32_t mul_row(const uint32_t (&in)[16], std::integer_sequence<int, J...>)
{
    uint32_t acc = 0u;
    ((acc ^= (MulBit<NBITS, I, J>::value ? in[J] : 0u)), ...);
    return acc;
}
template <int NBITS, int... I> OIP_BS_HD void mul_rows(const uint32_t (&in)[16], uint32_t (&out)[16], std::integer_sequence<int, I...>)
{
    ((out[I] = mul_row<NBITS, I>(in, std::make_integer_sequence<int, 16>{})), ...);
}
template <int NBITS> OIP_BS_HD void mul_xpow(const uint32_t (&in)[16], uint32_t (&out)[16])
{
    mul_rows<NBITS>(in, out, std::make_integer_sequence<int, 16>{});
}

// remainder of frame f out of the sliced registers
OIP_BS_HD uint32_t unslice(const uint32_t (&R)[16], int f)
{
    uint32_t c = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; ++i) c |= ((R[i] >> f) & 1u) << i;
    return c;
}

// One lane's piece: 7 words of each of the 32 frames.  load(j, T) fills T[f] with little-endian 32-bit word j of this
// lane's piece of frame f.  head = bytes in front of the message inside the span (even, < 28): lane 0 clears its
// state after them.
template <typename Load> OIP_BS_HD void piece32(Load load, int lane, int head, uint32_t (&P)[16])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; ++i) P[i] = 0u;
    const int clear_bit = lane == 0 ? 8 * head : -1; // bit position inside the piece where the state is cleared
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < 7; ++j) {
        uint32_t T[32];
        load(j, T);
        transpose32(T);
        lfsr_word(P, T, clear_bit - 32 * j);
    }
}

// butterfly join, distance s: xchg(v, s, i) = register i of lane ^ s.  The partner that covers the EARLIER bytes
// (lane & s == 0) is advanced by the s * UNIT bits of the later one (UNIT = bits of the message between the ends of
// two neighbouring lanes' data: 224 for contiguous 28-byte pieces, 32 for word-interleaved pieces).
template <int S, int UNIT = 8 * PIECE, typename Xchg> OIP_BS_HD void join_level(uint32_t (&P)[16], int lane, Xchg xchg)
{
    uint32_t M[16];
    mul_xpow<UNIT * S>(P, M);
    const bool early = (lane & S) == 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; ++i) {
        const uint32_t send = early ? M[i] : P[i];
        P[i] = send ^ xchg(send, S, i);
    }
}

// Word-interleaved pieces (coalesced loads straight from global memory): the 896-byte span is seven 128-byte blocks and
// lane l owns word l of every block, so a warp load instruction reads 128 contiguous bytes.  Between two of its words a
// lane's remainder is advanced by the 124 bytes in between (one fixed XOR network per block).  load(j, T) fills T[f]
// with word l of block j of frame f.  head = bytes in front of the message (even, < 128): words entirely inside it are
// dropped, the word that straddles it is entered from its upper half.
template <typename Load> OIP_BS_HD void piece32_interleaved(Load load, int lane, int head, uint32_t (&P)[16])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; ++i) P[i] = 0u;
    const int o = 4 * lane; // offset of this lane's word inside block 0
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < 7; ++j) {
        uint32_t T[32];
        load(j, T);
        transpose32(T);
        if (j > 0) {
            uint32_t Q[16];
            mul_xpow<8 * 124>(P, Q);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 16; ++i) P[i] = Q[i];
        }
        lfsr_word(P, T, (j == 0 && o < head && head < o + 4) ? 8 * (head - o) : -1);
        if (j == 0 && o + 4 <= head) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 16; ++i) P[i] = 0u;
        }
    }
}

#if defined(__CUDACC__)
// full warp: every lane returns the sliced remainders (zero initial value) of the 32 frames
template <typename Load> __device__ __forceinline__ void warp_crc32frames(Load load, int head, uint32_t (&P)[16])
{
    const int lane = threadIdx.x & 31;
    piece32(load, lane, head, P);
    auto x = [](uint32_t v, int s, int) { return __shfl_xor_sync(0xffffffffu, v, s); };
    join_level<1>(P, lane, x);
    join_level<2>(P, lane, x);
    join_level<4>(P, lane, x);
    join_level<8>(P, lane, x);
    join_level<16>(P, lane, x);
}
template <typename Load> __device__ __forceinline__ void warp_crc32frames_interleaved(Load load, int head, uint32_t (&P)[16])
{
    const int lane = threadIdx.x & 31;
    piece32_interleaved(load, lane, head, P);
    auto x = [](uint32_t v, int s, int) { return __shfl_xor_sync(0xffffffffu, v, s); };
    join_level<1, 32>(P, lane, x);
    join_level<2, 32>(P, lane, x);
    join_level<4, 32>(P, lane, x);
    join_level<8, 32>(P, lane, x);
    join_level<16, 32>(P, lane, x);
}
#endif

} // namespace bitslice
} // namespace oip
