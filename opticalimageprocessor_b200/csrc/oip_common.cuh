// oip_common.cuh -- shared plumbing of liboip_b200.so (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/oip_b200.h"

// one cached oip_pan_pipeline plan (device-resident tile lists), keyed on everything the split depends on
struct oip_pan_plan {
    std::vector<uint8_t> key;
    void *d_plan = nullptr;
    size_t cap = 0;
    int64_t tiles = 0, fast_ctas = 0;
    int64_t fast_cls[3] = {0, 0, 0}; // pan_fast_kernel CTAs per source-format class (line rasters / sub-image tiles / packed lines)
    size_t fast_off = 0;
    uint64_t last_use = 0;
    void *h_stage = nullptr;    // pinned staging of the plan upload (a pageable source would tie the host to the stream)
    size_t h_cap = 0;
    cudaEvent_t done = nullptr; // recorded after the last launch that read d_plan: a slot is recycled behind it
};

struct oip_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int64_t launches = 0;
    // cached device-side plan buffers (re-uploaded only when the geometry changes)
    // (the host-buffer pipeline calls oip_pan_pipeline once per row block: every block keeps its own plan)
    std::vector<oip_pan_plan> pan_plans;
    uint64_t plan_clock = 0;
    void *d_plan = nullptr;      // the plan of the current call (owned by pan_plans)
    int64_t plan_tiles = 0;      // generic tiles (pan_kernel CTAs)
    int64_t plan_fast_ctas = 0;  // pan_fast_kernel CTAs (4 warp-tiles each), all classes
    int64_t plan_fast_cls[3] = {0, 0, 0};
    size_t plan_fast_off = 0;    // byte offset of the FastTile array inside d_plan
    // tunables (oip_ctx_set_option)
    int host_block_rows = 2048;  // oip_pan_pipeline_host: rows per H2D / compute / D2H block
    int pan_fast = 1;            // 0: everything on the generic kernel
    int pan_fast_stages = 4;     // TMA stages per warp
    int pan_fast_rows = 128;     // output rows per warp-tile
    int pan_fast_minb = 3;       // register-allocation variant of pan_fast_kernel (CTAs per SM: 2, 3, 4)
    void *d_mss_plan = nullptr;
    size_t d_mss_plan_cap = 0;
    std::vector<uint8_t> mss_plan_key;
    int64_t mss_plan_tiles = 0;
    int64_t mss_plan_rows = 0;
    int64_t mss_fast_ctas = 0;   // mss_fast_kernel CTAs
    size_t mss_fast_off = 0;     // byte offset of the FTile array inside d_mss_plan
    int aos_fused = 1;           // 0: oip_aos_scan uses the exhaustive search kernels of round 1 (aos_scan_kernel + aos_crc_kernel)
    int downlink_threads = 1;    // 0: oip_downlink_to_stitched runs stage 1 of its CCDs one after the other on the caller's stream
    int imtr_runs = 1;           // 0: oip_imtr_deframe gathers frame by frame (imtr_validate_kernel) instead of run by run
    int mss_fast = 1;            // 0: every MSS tile on the generic kernel
    int mss_fast_rows = 128;     // output rows per MSS warp-tile
    // scratch for stage 1 (grown on demand)
    void *d_scratch = nullptr;
    size_t d_scratch_cap = 0;
    void *h_pinned = nullptr; // small pinned staging for counters / plans
    size_t h_pinned_cap = 0;
    int *d_err = nullptr;     // device-side error flag
    bool side_stream = true;  // generic tiles on aux_stream (off inside the host-buffer pipeline: its copy streams and the
                              // side stream can share a hardware queue, which parks the kernel behind a 2 ms copy)
    cudaStream_t aux_stream = nullptr; // side stream of oip_pan_pipeline (generic tiles next to the fast kernel)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool pan_attr_set = false, mss_attr_set = false, fast_attr_set[3] = {false, false, false}, imtr_attr_set = false;
    // host-buffer pipeline (oip_pan_pipeline_host): staging slots + side streams
    void *host_pipe = nullptr;
    void *stt_state = nullptr; // cuFFT plans of the offset estimation (stt.cu)
    void *downlink_state = nullptr; // per-CCD IMDT / table buffers of oip_downlink_to_stitched (downlink.cu)
};

namespace oip {

void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#define OIP_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return oip::fail(OIP_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                           \
    } while (0)

#define OIP_CHECK_CTX(ctx)                                                     \
    do {                                                                       \
        if (!(ctx)) return oip::fail(OIP_E_INVALID, "null context");           \
        OIP_CUDA(cudaSetDevice((ctx)->device));                                \
    } while (0)

int ensure_scratch(oip_ctx *ctx, size_t bytes);
int ensure_pinned(oip_ctx *ctx, size_t bytes);
void host_pipe_destroy(oip_ctx *ctx);
namespace stt { void destroy(oip_ctx *ctx); }
void downlink_destroy(oip_ctx *ctx);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA unit, SASS UBLKCP); 16-byte aligned src/dst/size
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t smem_addr)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(smem_addr));
    return r;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_na_v4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
// packed fp32 pairs: two pixels per FFMA2/FADD2, each lane rounds exactly like the scalar op
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f2 v)
{
    float a;
    asm("{\n\t.reg .f32 t;\n\tmov.b64 {%0, t}, %1;\n\t}" : "=f"(a) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi_of(f2 v)
{
    float b;
    asm("{\n\t.reg .f32 t;\n\tmov.b64 {t, %0}, %1;\n\t}" : "=f"(b) : "l"(v));
    return b;
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false (it honours the
// explicit .rn only for scalar ops), and it also folds fma(a,b,-0.0) back into a mul when the -0.0 is
// a known constant.  The product is therefore an FMA whose addend is a (-0.0,-0.0) pair LOADED AT RUN
// TIME (plan header): RN(a*b + -0.0) == RN(a*b) bit for bit, and an FMA cannot be fused with the add
// that follows.  SASS check: FFMA2 count == number of products, FADD2 count == number of sums.
__device__ __forceinline__ f2 mul2(f2 a, f2 b, f2 nz)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint32_t bswap16x2(uint32_t v) { return __byte_perm(v, 0, 0x2301); }

// CRC-16/CCITT-FALSE one byte step (poly 0x1021, MSB first) -- same recurrence as ref CRC.h:806-834
__device__ __forceinline__ uint32_t crc16_byte(uint32_t crc, uint32_t byte)
{
    uint32_t x = ((crc >> 8) ^ byte) & 0xFFu;
    x ^= x >> 4;
    return ((crc << 8) ^ (x << 12) ^ (x << 5) ^ x) & 0xFFFFu;
}

// (uint16_t)(k*s + b) exactly as the reference's x86 build evaluates it (ref imageop.h:134;
// cvtsi2sd, mulsd, addsd, cvttsd2si(32-bit), low 16 bits).  No FMA contraction.
// The conversion pipe (I2F/F2I .F64) issues at 16 lanes/clk/SM on B200 against 64 for DADD/DMUL
// (tools/probes/pipe_rates.cu), so both conversions are done with exact magic-number adds instead:
//   u16 -> f64 : (2^52 | s) - 2^52                      (exact, one DADD)
//   trunc      : low word of RZ(v + 2^52) for 0 <= v < 2^31 (exact, one DADD.RZ); anything else
//                takes the generic path (negative, >= 2^31, NaN: rare, input-contract territory)
__device__ __forceinline__ uint32_t rrc_px(uint32_t s, double k, double b)
{
    const double sd = __dadd_rn(__hiloint2double(0x43300000, (int)s), -4503599627370496.0);
    const double v = __dadd_rn(__dmul_rn(k, sd), b);
    const uint32_t hi = (uint32_t)__double2hiint(v);
    if (hi < 0x41E00000u) // sign clear and exponent < 1023+31  <=>  0 <= v < 2^31
        return (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0)) & 0xFFFFu;
    int t = (v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : (int)0x80000000;
    return (uint32_t)t & 0xFFFFu;
}
// exact u16 -> f32 without the conversion pipe
__device__ __forceinline__ float u16_to_f32(uint32_t v)
{
    return __fadd_rn(__uint_as_float(0x4B000000u | v), -8388608.0f);
}
#endif

} // namespace oip
