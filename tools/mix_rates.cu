// mix_rates.cu -- does a packed FP32 instruction (FFMA2) leave the issue slot of its second pipe cycle to other pipes?
// Prints cycles per loop iteration per SM sub-partition for mixes of independent chains (8 warps per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int NP, int NS, int NL, int ND, int NX>
__global__ void k(float *out, int n, float a0, double d0, int i0)
{
    unsigned long long p2[NP ? NP : 1]; float f[NS ? NS : 1]; int ii[NL ? NL : 1]; double d[ND ? ND : 1]; float x[NX ? NX : 1];
    for (int c = 0; c < NP; ++c) p2[c] = (unsigned long long)__float_as_uint(a0 + c) << 32 | __float_as_uint(a0 + c + threadIdx.x);
    for (int c = 0; c < NS; ++c) f[c] = a0 + c + threadIdx.x;
    for (int c = 0; c < NL; ++c) ii[c] = i0 + c + threadIdx.x;
    for (int c = 0; c < ND; ++c) d[c] = d0 + c + threadIdx.x;
    for (int c = 0; c < NX; ++c) x[c] = a0 + c + threadIdx.x;
    unsigned long long w2 = (unsigned long long)__float_as_uint(a0) << 32 | __float_as_uint(a0);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (c < NP) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2[c]) : "l"(w2));
            if (c < NS) f[c] = __fmaf_rn(f[c], a0, a0);
            if (c < NL) ii[c] = (ii[c] ^ i0) & (ii[c] | it);
            if (c < ND) d[c] = __dadd_rn(d[c], d0);
            if (c < NX) { unsigned short r; asm volatile("cvt.rni.u16.f32 %0, %1;" : "=h"(r) : "f"(x[c])); x[c] = __int_as_float((int)r + i0); }
        }
    }
    float s = 0;
    for (int c = 0; c < NP; ++c) s += (float)(p2[c] >> 40);
    for (int c = 0; c < NS; ++c) s += f[c];
    for (int c = 0; c < NL; ++c) s += ii[c];
    for (int c = 0; c < ND; ++c) s += (float)d[c];
    for (int c = 0; c < NX; ++c) s += x[c];
    if (s == 12345.678f) out[0] = s;
}
template <int NP, int NS, int NL, int ND, int NX> void run(const char *name)
{
    float *out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount, threads = 1024; // 32 warps per SM = 8 per SMSP
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NP, NS, NL, ND, NX><<<blocks, threads>>>(out, 64, 1.0001f, 1.0001, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<NP, NS, NL, ND, NX><<<blocks, threads>>>(out, ITER, 1.0001f, 1.0001, 3);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3 / ITER / 8.0; // cycles per iteration per warp-slot (8 warps share an SMSP)
    printf("%-44s %7.3f ms  %6.2f cycles per iteration per warp (issue slots needed: %d)\n", name, ms, cyc, NP + NS + NL + ND + NX);
    cudaFree(out);
}
int main()
{
    run<8, 0, 0, 0, 0>("8 FFMA2");
    run<0, 16, 0, 0, 0>("16 FFMA (same flops)");
    run<8, 0, 8, 0, 0>("8 FFMA2 + 8 LOP3");
    run<8, 0, 16, 0, 0>("8 FFMA2 + 16 LOP3");
    run<0, 16, 8, 0, 0>("16 FFMA + 8 LOP3");
    run<8, 0, 0, 4, 0>("8 FFMA2 + 4 DADD");
    run<8, 0, 0, 8, 0>("8 FFMA2 + 8 DADD");
    run<8, 0, 0, 0, 2>("8 FFMA2 + 2 F2I.U16(+IADD)");
    run<8, 0, 4, 2, 1>("8 FFMA2 + 4 LOP3 + 2 DADD + 1 F2I(+IADD)");
    run<0, 0, 16, 0, 0>("16 LOP3");
    run<0, 0, 0, 8, 0>("8 DADD");
    return 0;
}
