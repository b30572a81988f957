// mix_rates.cu -- what does one instruction of class X cost when it runs next to a packed-FP32 stream?
// Each kernel iteration issues 8 FFMA2 (independent chains) plus N instructions of one other class; 8 warps per SMSP.
// Output: extra SMSP cycles per added warp-instruction (0 = fully hidden behind the FP32 pipe, 2.2 = as expensive as an FFMA2).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
enum { NONE, LOP3, PRMT, MOVI, IADD, IMAD, DADD, DFMA, I2F16, I2F64, F2I16, SHFL, LDS32, LDS64, FADDS, ISETP };
template <int NP, int X, int NX>
__global__ void k(float *out, int n, float a0, double d0, int i0)
{
    __shared__ float sm[1024];
    unsigned long long p2[NP ? NP : 1];
    int ii[NX ? NX : 1]; double d[NX ? NX : 1]; float x[NX ? NX : 1];
    for (int c = 0; c < NP; ++c) p2[c] = (unsigned long long)__float_as_uint(a0 + c) << 32 | __float_as_uint(a0 + c + threadIdx.x);
    for (int c = 0; c < NX; ++c) { ii[c] = i0 + c + threadIdx.x; d[c] = d0 + c + threadIdx.x; x[c] = a0 + c + threadIdx.x; }
    sm[threadIdx.x] = a0;
    __syncthreads();
    unsigned long long w2 = (unsigned long long)__float_as_uint(a0) << 32 | __float_as_uint(a0);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (c < NP) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2[c]) : "l"(w2));
            if (c < NX) {
                if (X == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ii[c]) : "r"(i0), "r"(it));
                if (X == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2301;" : "+r"(ii[c]) : "r"(i0));
                if (X == MOVI) asm volatile("mov.b32 %0, %1;" : "=r"(ii[c]) : "r"(it + c));
                if (X == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(ii[c]) : "r"(i0));
                if (X == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ii[c]) : "r"(i0), "r"(it));
                if (X == DADD) d[c] = __dadd_rn(d[c], d0);
                if (X == DFMA) d[c] = __fma_rn(d[c], d0, d0);
                if (X == I2F16) { unsigned short h = (unsigned short)ii[c]; asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(x[c]) : "h"(h)); ii[c] = __float_as_int(x[c]); }
                if (X == I2F64) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[c]) : "r"(ii[c])); ii[c] = __double2hiint(d[c]); }
                if (X == F2I16) { unsigned short r; asm volatile("cvt.rni.u16.f32 %0, %1;" : "=h"(r) : "f"(x[c])); x[c] = __int_as_float((int)r | 0x3f800000); }
                if (X == SHFL) ii[c] = __shfl_down_sync(0xffffffffu, ii[c], 1);
                if (X == LDS32) ii[c] = __float_as_int(sm[(ii[c] + threadIdx.x) & 1023]);
                if (X == LDS64) { float2 v = *reinterpret_cast<float2 *>(&sm[((ii[c] + threadIdx.x) * 2) & 1022]); ii[c] = __float_as_int(v.x) + __float_as_int(v.y); }
                if (X == FADDS) x[c] = __fadd_rn(x[c], a0);
                if (X == ISETP) { if (ii[c] > it) ii[c] = it; }
            }
        }
    }
    float s = 0;
    for (int c = 0; c < NP; ++c) s += (float)(p2[c] >> 40);
    for (int c = 0; c < NX; ++c) s += ii[c] + (float)d[c] + x[c];
    if (s == 12345.678f) out[0] = s;
}
static float base_cycles = 0;
template <int NP, int X, int NX> void run(const char *name)
{
    float *out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount, threads = 1024;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NP, X, NX><<<blocks, threads>>>(out, 64, 1.0001f, 1.0001, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<NP, X, NX><<<blocks, threads>>>(out, ITER, 1.0001f, 1.0001, 3);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3 / ITER; // SMSP cycles per iteration (all 8 warps)
    if (NX == 0) base_cycles = (float)cyc;
    printf("%-28s %8.1f SMSP cycles/iter", name, cyc);
    if (NX) printf("   extra per added warp-instr: %5.2f cycles", (cyc - (NP ? base_cycles : 0)) / (8.0 * NX));
    printf("\n");
    cudaFree(out);
}
int main()
{
    printf("--- issue-port cost: 16 FFMA2 + 2 X per warp iteration (no pipe but the FP32 one near saturation)\n");
    run<16, NONE, 0>("16 FFMA2 (x8 warps)");
#define PORT(X) run<16, X, 2>("16 FFMA2 + 2 " #X);
    PORT(LOP3) PORT(PRMT) PORT(MOVI) PORT(IADD) PORT(IMAD) PORT(ISETP) PORT(FADDS) PORT(DADD) PORT(DFMA) PORT(I2F16) PORT(I2F64) PORT(F2I16) PORT(SHFL) PORT(LDS32) PORT(LDS64)
    printf("--- 1:1 mixes\n");
    run<8, NONE, 0>("8 FFMA2 (x8 warps)");
#define MIX(X) run<8, X, 8>("8 FFMA2 + 8 " #X); run<0, X, 8>("          8 " #X " alone");
    MIX(LOP3) MIX(PRMT) MIX(MOVI) MIX(IADD) MIX(IMAD) MIX(ISETP) MIX(FADDS) MIX(DADD) MIX(DFMA) MIX(I2F16) MIX(I2F64) MIX(F2I16) MIX(SHFL) MIX(LDS32) MIX(LDS64)
    return 0;
}
