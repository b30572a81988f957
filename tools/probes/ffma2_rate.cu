// ffma2_rate.cu -- issue rate of the packed FP32 forms the bicubic kernels use, per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o /tmp/ffma2_rate tools/probes/ffma2_rate.cu && /tmp/ffma2_rate
// Variants: 0 FFMA2 pair,scalar,pair(R)   1 FFMA2 pair,scalar,pair(UR)   2 FMUL2 pair,scalar   3 FADD2 pair,pair
//           4 FFMA2 pair,pair,pair(R)     5 FFMA2 pair,pair,pair(UR)
//           6 [FFMA2(R addend) ; FADD2 dependent] pairs as in the kernel   7 same with the UR addend
//           8 FFMA (scalar fp32, 3 registers)    9 DFMA
//           10 FADD2 + LOP3 interleaved 1:1   11 FADD2 + DFMA 2:1   12 FMUL scalar   13 FADD scalar
//           14 FFMA2 pair,scalar,scalar(-0 broadcast)   15 FFMA2(UR) + LOP3 1:1   16 FMUL+FADD scalar + LOP3 (2:1)
//           17 FADD2 + I2F.U16 2:1   18 LOP3 alone   19 FADD2 + SHFL 4:1   20 FADD2+DFMA+LOP3 4:2:2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
struct P { f2 nz; const f2 *in; f2 *out; long long *cyc; int iters; };
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
constexpr int NA = 12;
template <int V> __global__ void __launch_bounds__(512, 1) k(const __grid_constant__ P p)
{
    f2 x[NA];
    const int t = threadIdx.x;
    for (int i = 0; i < NA; ++i) x[i] = p.in[(t + i) & 255];
    const float w0 = __uint_as_float((uint32_t)(p.in[3] & 0xffffffffu)), w1 = __uint_as_float((uint32_t)(p.in[5] & 0xffffffffu));
    const f2 ws0 = pk(w0, w0), ws1 = pk(w1, w1), wp0 = p.in[7 + (t & 1)], wp1 = p.in[9 + (t & 1)];
    f2 nzr = p.in[11];               // vector-register copy of the addend
    const f2 nzu = p.nz;             // kernel parameter: uniform register / constant bank
    double d[NA];
    for (int i = 0; i < NA; ++i) d[i] = (double)(t + i);
    float s[NA], s2[NA];
    for (int i = 0; i < NA; ++i) { s[i] = (float)(t + i); s2[i] = (float)(t - i); }
    uint32_t y[NA];
    for (int i = 0; i < NA; ++i) y[i] = (uint32_t)(p.in[20 + i] >> 7) + t;
    const uint32_t c0 = (uint32_t)p.in[40];
    const float zneg = __uint_as_float((uint32_t)(p.in[11] >> 32));
    const f2 nzb = pk(zneg, zneg);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (V == 0) x[i] = fma2(x[i], (i & 1) ? ws1 : ws0, nzr);
            if (V == 1) x[i] = fma2(x[i], (i & 1) ? ws1 : ws0, nzu);
            if (V == 2) x[i] = mul2(x[i], (i & 1) ? ws1 : ws0);
            if (V == 3) x[i] = add2(x[i], (i & 1) ? wp1 : wp0);
            if (V == 4) x[i] = fma2(x[i], (i & 1) ? wp1 : wp0, nzr);
            if (V == 5) x[i] = fma2(x[i], (i & 1) ? wp1 : wp0, nzu);
            if (V == 6) x[i] = add2(x[i], fma2(x[(i + 1) % NA], (i & 1) ? ws1 : ws0, nzr));
            if (V == 7) x[i] = add2(x[i], fma2(x[(i + 1) % NA], (i & 1) ? ws1 : ws0, nzu));
            if (V == 8) s[i] = __fmaf_rn(s[i], w0, w1);
            if (V == 9) d[i] = __fma_rn(d[i], 1.0000001, 0.5);
            if (V == 10) { x[i] = add2(x[i], (i & 1) ? wp1 : wp0); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 5) % NA]), "r"(c0)); }
            if (V == 11) { x[i] = add2(x[i], (i & 1) ? wp1 : wp0); if (i & 1) d[i] = __fma_rn(d[i], 1.0000001, 0.5); }
            if (V == 12) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(w0));
            if (V == 13) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(w0));
            if (V == 14) x[i] = fma2(x[i], (i & 1) ? ws1 : ws0, nzb);
            if (V == 15) { x[i] = fma2(x[i], (i & 1) ? ws1 : ws0, nzu); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 5) % NA]), "r"(c0)); }
            if (V == 16) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(w0)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s2[i]) : "f"(w1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 5) % NA]), "r"(c0)); }
            if (V == 17) { x[i] = add2(x[i], (i & 1) ? wp1 : wp0); if (i & 1) { uint32_t u = y[i] & 0xffffu; asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(s[i]) : "h"((unsigned short)u)); } }
            if (V == 18) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 5) % NA]), "r"(c0));
            if (V == 19) { x[i] = add2(x[i], (i & 1) ? wp1 : wp0); if ((i & 3) == 0) y[i] = __shfl_down_sync(0xffffffffu, y[i], 1); }
            if (V == 20) { x[i] = add2(x[i], (i & 1) ? wp1 : wp0); if (i & 1) d[i] = __fma_rn(d[i], 1.0000001, 0.5);
                           else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 5) % NA]), "r"(c0)); }
        }
    }
    const long long t1 = clock64();
    f2 acc = 0;
    for (int i = 0; i < NA; ++i) acc ^= x[i] ^ (f2)__double_as_longlong(d[i]) ^ (f2)__float_as_uint(s[i]) ^ (f2)__float_as_uint(s2[i]) ^ ((f2)y[i] << 13);
    p.out[blockIdx.x * blockDim.x + t] = acc;
    if ((t & 31) == 0) p.cyc[blockIdx.x * 16 + (t >> 5)] = t1 - t0;
}
template <int V> static void run(const P &p0, int warps, const char *name)
{
    P p = p0;
    k<V><<<148, warps * 32>>>(p);
    cudaDeviceSynchronize();
    k<V><<<148, warps * 32>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, p.cyc, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    const double iter_per_smsp = (double)p.iters * (warps / 4.0);
    printf("%-44s warps/SMSP %d  cycles per 12-slot group per SMSP %.2f  %s\n", name, warps / 4, mx / iter_per_smsp, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main()
{
    P p{};
    p.nz = 0x8000000080000000ull;
    p.iters = 4096;
    f2 h[256];
    for (int i = 0; i < 256; ++i) { float a = 1.0f + i * 1e-6f, b = 1.0f - i * 1e-6f; uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); h[i] = ((f2)ub << 32) | ua; }
    h[11] = p.nz;
    cudaMalloc(&p.in, sizeof h); cudaMemcpy((void *)p.in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMalloc(&p.out, 148 * 512 * 8); cudaMalloc(&p.cyc, 148 * 16 * 8);
    for (int w : {4, 8, 12, 16}) {
        if (w == 4) run<0>(p, 4, "FFMA2 pair,scalar,pair(R)"); if (w == 8) run<0>(p, 8, "FFMA2 pair,scalar,pair(R)"); if (w == 12) run<0>(p, 12, "FFMA2 pair,scalar,pair(R)"); if (w == 16) run<0>(p, 16, "FFMA2 pair,scalar,pair(R)");
    }
#define ALLW(V, NAME) run<V>(p, 4, NAME); run<V>(p, 8, NAME); run<V>(p, 12, NAME); run<V>(p, 16, NAME);
    ALLW(1, "FFMA2 pair,scalar,pair(UR)")
    ALLW(2, "FMUL2 pair,scalar")
    ALLW(3, "FADD2 pair,pair")
    ALLW(4, "FFMA2 pair,pair,pair(R)")
    ALLW(5, "FFMA2 pair,pair,pair(UR)")
    ALLW(6, "FFMA2(R addend)+FADD2 dependent pairs")
    ALLW(7, "FFMA2(UR addend)+FADD2 dependent pairs")
    ALLW(8, "FFMA scalar")
    ALLW(9, "DFMA")
    ALLW(10, "12 FADD2 + 12 LOP3")
    ALLW(11, "12 FADD2 + 6 DFMA")
    ALLW(12, "12 FMUL scalar")
    ALLW(13, "12 FADD scalar")
    ALLW(14, "12 FFMA2 pair,scalar,bcast(-0)")
    ALLW(15, "12 FFMA2(UR) + 12 LOP3")
    ALLW(16, "12 FMUL + 12 FADD scalar + 12 LOP3")
    ALLW(17, "12 FADD2 + 6 I2F.U16")
    ALLW(18, "12 LOP3")
    ALLW(19, "12 FADD2 + 3 SHFL")
    ALLW(20, "12 FADD2 + 6 DFMA + 6 LOP3")
    return 0;
}
