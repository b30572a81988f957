// probe: which 2-D TMA tensor-load forms run on this box?  usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// variant: 0 = 1-D bulk copy (sanity), 1 = tensor.2d plain, 2 = tensor.2d with L2 cache hint, 3 = tensor.2d shared::cta,
//          4 = plain after prefetch.tensormap, 5 = plain + fence.proxy.tensormap acquire
__global__ void k(const __grid_constant__ CUtensorMap tmp, const CUtensorMap *tmg, const uint16_t *src, uint16_t *out, int variant, int bytes,
                  int use_param, int cx, int cy)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const CUtensorMap *tm = use_param ? &tmp : tmg;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        if (variant == 4) asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
        if (variant == 5) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes));
        if (variant == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm)), "l"(src), "r"(bytes),
                         "r"(s32(&bar)) : "memory");
        else if (variant == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(s32(sm)),
                         "l"(tm), "r"(0), "r"(0), "r"(s32(&bar)), "l"(0x1000000000000000ull) : "memory");
        else if (variant == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                         "l"(tm), "r"(0), "r"(0), "r"(s32(&bar)) : "memory");
        else if (variant == 6)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                         "l"(tm), "r"(cx), "r"(cy), "r"(s32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                         "l"(tm), "r"(cx), "r"(cy), "r"(s32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    int spins = 0;
    while (!ok && spins < (1 << 22)) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
        ++spins;
    }
    for (int i = threadIdx.x; i < bytes / 2; i += 32) out[i] = ok ? ((uint16_t *)sm)[i] : 0xDEAD;
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    int variant = argc > 1 ? atoi(argv[1]) : 1, use_param = argc > 2 ? atoi(argv[2]) : 1, box_w = argc > 3 ? atoi(argv[3]) : 64;
    int cx = argc > 4 ? atoi(argv[4]) : 0, cy = argc > 5 ? atoi(argv[5]) : 0, promo = argc > 6 ? atoi(argv[6]) : 0;
    const int W = 1024, H = 64;
    std::vector<uint16_t> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (uint16_t)(i * 7 + 1);
    uint16_t *d, *o;
    cudaMalloc(&d, W * H * 2); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), W * H * 2, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Enc enc = (Enc)p;
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    cuuint64_t dims[2] = {W, H}, str[1] = {W * 2}; cuuint32_t box[2] = {(cuuint32_t)box_w, 4}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const uint64_t *w64 = (const uint64_t *)&tm;
    printf("variant %d param %d box_w %d xy %d,%d promo %d: entry=%d q=%d encode=%d map[0..3]=%016llx %016llx %016llx %016llx\n", variant, use_param, box_w, cx, cy, promo, (int)ge, (int)q, (int)r,
           (unsigned long long)w64[0], (unsigned long long)w64[1], (unsigned long long)w64[2], (unsigned long long)w64[3]);
    CUtensorMap *dt; cudaMalloc(&dt, 128); cudaMemcpy(dt, &tm, 128, cudaMemcpyHostToDevice);
    int bytes = variant == 0 ? 2048 : box_w * 2 * 4;
    k<<<1, 32, 16384>>>(tm, dt, d, o, variant, bytes, use_param, cx, cy);
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> g(bytes / 2);
    cudaMemcpy(g.data(), o, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (variant == 0) for (int i = 0; i < bytes / 2; ++i) bad += g[i] != h[i];
    else for (int rr = 0; rr < 4; ++rr) for (int c = 0; c < box_w; ++c) bad += g[rr * box_w + c] != h[(rr + cy) * W + c + cx];
    printf("   launch: %s, sync: %s, mismatches %d, first %04x\n", cudaGetErrorString(e1), cudaGetErrorString(e), bad, g[0]);
    return 0;
}
