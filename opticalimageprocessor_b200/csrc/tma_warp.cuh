// tma_warp.cuh -- device helpers shared by the warp-tile kernels (pan_fast.cu, mss_fast.cu): 2-D TMA tensor loads into
// per-warp stage rings behind mbarriers, shared-memory loads, exact sample conversion on the FP64 / conversion pipes.
#pragma once
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "oip_common.cuh"

namespace oip {
namespace tmaw {

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
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }
__device__ __forceinline__ f2 shfl_down1(f2 v)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_down_sync(0xffffffffu, lo, 1);
    hi = __shfl_down_sync(0xffffffffu, hi, 1);
    return ((f2)hi << 32) | lo;
}

// On B200 the integer/logic instructions (PRMT, LOP3, MOV, IADD3 ...) take their cycles from the same datapath
// as the FP32 instructions (tools/probes/mix_rates.cu: FFMA2 + LOP3 times add up, FFMA2 + DADD overlap), and the FP32
// datapath is what bounds this kernel.  So a 32-bit word of two samples is turned into two exact doubles on the
// conversion and FP64 pipes alone: I2F.F64.U32, then hi = RZ(x*2^-16 + 2^52) - 2^52, lo = x - 65536*hi.
struct D2 { double lo, hi; };
__device__ __forceinline__ D2 split_word(uint32_t w)
{
    const double x = __uint2double_rn(w); // I2F.F64.U32: one issue slot (the magic-number form costs MOV + DADD)
    D2 r;
    r.hi = __dadd_rn(__fma_rz(x, 1.52587890625e-05, 4503599627370496.0), -4503599627370496.0);
    r.lo = __fma_rn(r.hi, -65536.0, x);
    return r;
}
// RRC of an exact sample value.  MODE 1: every (k,b) of the warp is >= 0 and k*65535+b < 2^31: truncation is the
// low word of RZ(v + 2^52).  MODE 2: general (sign / range handling exactly like x86 cvttsd2si).  The low 16
// bits of the result are taken by the consumer (I2F.U16 / PRMT).
template <int MODE>
__device__ __forceinline__ uint32_t rrc_d(double sd, double k, double b)
{
    const double v = __dadd_rn(__dmul_rn(k, sd), b);
    if (MODE == 1) return (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0));
    const uint32_t hi = (uint32_t)__double2hiint(v);
    if (hi < 0x41E00000u) return (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0));
    return (uint32_t)((v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : (int)0x80000000);
}

// host: encode a 2-D tiled tensor map over rows of `w` u16 samples counted as 32-bit elements (box_w32 x box_rows)
int encode_tmap_u32(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes, int box_w32, int box_rows);

} // namespace tmaw
} // namespace oip
