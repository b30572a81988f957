// dot_rates.cu -- cycles per source row of the bicubic scatter-form inner product alone (no loads, no conversion):
// the FP32-datapath floor of pan_fast_kernel.  Same instruction stream as remap_tile's row body: 64 FFMA2 + 60 FADD2.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b, f2 nz) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int WARPS_PER_SMSP>
__global__ void __launch_bounds__(WARPS_PER_SMSP * 128) k(float *out, const float *tab, int rows, float seed)
{
    const f2 nz = *reinterpret_cast<const f2 *>(tab + 16);
    f2 W[4][4];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { float w = tab[4 * r + c]; W[r][c] = pk(w, w); }
    f2 B0[4], B1[4], B2[4], B3[4], win[7];
    for (int o = 0; o < 4; ++o) B0[o] = B1[o] = B2[o] = B3[o] = 0ull;
    for (int j = 0; j < 7; ++j) win[j] = pk(seed + j + threadIdx.x, seed - j);
    f2 acc = 0;
    for (int m = 0; m < rows; m += 4) {
        auto row = [&](f2(&AN)[4], f2(&A1)[4], f2(&A2)[4], f2(&A3)[4]) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                auto dot = [&](const f2(&Wr)[4]) {
                    return add2(add2(add2(mul2(win[o], Wr[0], nz), mul2(win[o + 1], Wr[1], nz)), mul2(win[o + 2], Wr[2], nz)), mul2(win[o + 3], Wr[3], nz));
                };
                AN[o] = dot(W[0]);
                A1[o] = add2(A1[o], dot(W[1]));
                A2[o] = add2(A2[o], dot(W[2]));
                f2 outv = add2(A3[o], dot(W[3]));
                acc ^= outv; // one LOP3 pair per output instead of the store
            }
#pragma unroll
            for (int j = 0; j < 7; ++j) win[j] ^= acc & 0x0000000100000001ull; // keep the window data-dependent on the loop
        };
        row(B0, B1, B2, B3); row(B3, B0, B1, B2); row(B2, B3, B0, B1); row(B1, B2, B3, B0);
    }
    if (acc == 0x1234567ull) out[0] = 1.f;
}
template <int WPS> void run()
{
    float *out, *tab; cudaMalloc(&out, 4); cudaMalloc(&tab, 128);
    float h[18]; for (int i = 0; i < 16; ++i) h[i] = 0.01f * (i + 1); h[16] = h[17] = -0.0f;
    cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int rows = 16384;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<WPS><<<p.multiProcessorCount, WPS * 128>>>(out, tab, 64, 1.f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<WPS><<<p.multiProcessorCount, WPS * 128>>>(out, tab, rows, 1.f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%d warps/SMSP: %.3f ms, %.1f cycles per warp-row per SMSP (124 packed instr; 2 cycles each = 248) err=%s\n", WPS, ms,
           ms * 1e-3 * clk * 1e3 / rows / WPS, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<1>(); run<2>(); run<3>(); run<4>(); return 0; }
