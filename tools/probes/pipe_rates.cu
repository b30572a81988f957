// pipe_rates.cu -- instruction throughput probe for sm_100a (lane-ops per clock per SM).
// Used to decide which pipe bounds the fused PAN kernel (DESIGN.md "pipe budget").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define CHAINS 8
template <int OP> __global__ void k(float *out, int n, float a0, double d0, int i0)
{
    float f[CHAINS]; double d[CHAINS]; int ii[CHAINS]; unsigned long long p2[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { f[c] = a0 + c + threadIdx.x; d[c] = d0 + c + threadIdx.x; ii[c] = i0 + c + threadIdx.x; p2[c] = (unsigned long long)__float_as_uint(f[c]) << 32 | __float_as_uint(f[c] + 1.f); }
    unsigned long long w2 = (unsigned long long)__float_as_uint(a0) << 32 | __float_as_uint(a0);
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) f[c] = __fmul_rn(f[c], a0);
            if (OP == 1) f[c] = __fadd_rn(f[c], a0);
            if (OP == 2) f[c] = __fmaf_rn(f[c], a0, a0);
            if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p2[c]) : "l"(w2));
            if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2[c]) : "l"(w2));
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2[c]) : "l"(w2));
            if (OP == 6) ii[c] = (ii[c] ^ i0) & (ii[c] | it);          // LOP3
            if (OP == 7) ii[c] = __funnelshift_l(ii[c], i0, 5);        // SHF
            if (OP == 8) ii[c] = __byte_perm(ii[c], i0, 0x2301);       // PRMT
            if (OP == 9) ii[c] = ii[c] * i0 + it;                      // IMAD
            if (OP == 10) d[c] = __dadd_rn(d[c], d0);
            if (OP == 11) d[c] = __dmul_rn(d[c], d0);
            if (OP == 12) d[c] = __fma_rn(d[c], d0, d0);
            if (OP == 13) { d[c] = (double)(unsigned)ii[c]; ii[c] += (int)__double2hiint(d[c]); } // I2F.F64 (+IADD)
            if (OP == 14) { ii[c] = __double2int_rz(d[c]); d[c] = __longlong_as_double(__double_as_longlong(d[c]) + ii[c]); } // F2I.F64
            if (OP == 15) { f[c] = (float)(unsigned)ii[c]; ii[c] += __float_as_int(f[c]); }     // I2F.F32
            if (OP == 16) { ii[c] = __float2int_rn(f[c]); f[c] = __int_as_float(__float_as_int(f[c]) + ii[c]); } // F2I.F32
            if (OP == 17) { f[c] = sm[(__float_as_int(f[c]) + threadIdx.x) & 4095]; }              // LDS.32 dependent
            if (OP == 18) { float4 v = *reinterpret_cast<float4 *>(&sm[((__float_as_int(f[c]) + threadIdx.x) * 4) & 4092]); f[c] = v.x + v.w; }
            if (OP == 19) { f[c] = __double2float_rn(d[c]); d[c] = __longlong_as_double(__double_as_longlong(d[c]) + __float_as_int(f[c])); } // F2F.F32.F64
        }
    }
    float s = 0; for (int c = 0; c < CHAINS; ++c) s += f[c] + (float)d[c] + ii[c] + (float)(p2[c] >> 40);
    if (s == 12345.678f) out[0] = s;
}
template <int OP> void run(const char *name, int lanes_per_op)
{
    float *out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 2, threads = 1024;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<blocks, threads, 16384>>>(out, 64, 1.0001f, 1.0001, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<OP><<<blocks, threads, 16384>>>(out, ITER, 1.0001f, 1.0001, 3);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * ITER * CHAINS * lanes_per_op;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-14s %8.3f ms  %8.1f Glane-op/s  %6.1f lane-op/clk/SM @%d MHz(max) err=%s\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main()
{
    run<0>("FMUL", 1); run<1>("FADD", 1); run<2>("FFMA", 1); run<3>("MUL.F32x2", 2); run<4>("ADD.F32x2", 2); run<5>("FMA.F32x2", 2);
    run<6>("LOP3x2", 2); run<7>("SHF", 1); run<8>("PRMT", 1); run<9>("IMAD", 1);
    run<10>("DADD", 1); run<11>("DMUL", 1); run<12>("DFMA", 1); run<13>("I2F.F64(+2)", 1); run<14>("F2I.F64(+2)", 1);
    run<15>("I2F.F32(+1)", 1); run<16>("F2I.F32(+1)", 1); run<17>("LDS.32", 1); run<18>("LDS.128", 1); run<19>("F2F.32.64(+2)", 1);
    return 0;
}
