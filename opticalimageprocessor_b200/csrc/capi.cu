// capi.cu -- context, error channel and memory helpers of the C ABI (include/oip_b200.h).
#include "oip_common.cuh"

namespace oip {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int ensure_scratch(oip_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->d_scratch_cap) return OIP_OK;
    if (ctx->d_scratch) {
        OIP_CUDA(cudaStreamSynchronize(ctx->stream));
        OIP_CUDA(cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
        ctx->d_scratch_cap = 0;
    }
    size_t cap = bytes + bytes / 4 + (1 << 20);
    OIP_CUDA(cudaMalloc(&ctx->d_scratch, cap));
    ctx->d_scratch_cap = cap;
    return OIP_OK;
}

int ensure_pinned(oip_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_pinned_cap) return OIP_OK;
    if (ctx->h_pinned) {
        OIP_CUDA(cudaStreamSynchronize(ctx->stream));
        OIP_CUDA(cudaFreeHost(ctx->h_pinned));
        ctx->h_pinned = nullptr;
        ctx->h_pinned_cap = 0;
    }
    size_t cap = bytes * 2 + 4096;
    OIP_CUDA(cudaMallocHost(&ctx->h_pinned, cap));
    ctx->h_pinned_cap = cap;
    return OIP_OK;
}

} // namespace oip

extern "C" {

const char *oip_last_error(void) { return oip::g_err; }
int oip_abi_version(void) { return OIP_ABI_VERSION; }

int oip_ctx_create(int device, void *stream, int own_stream, oip_ctx **out)
{
    if (!out) return oip::fail(OIP_E_INVALID, "oip_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return oip::fail(OIP_E_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return oip::fail(OIP_E_INVALID, "device %d out of range (%d devices)", device, n);
    OIP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OIP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return oip::fail(OIP_E_CUDA, "device %d is sm_%d%d; liboip_b200 carries sm_100a code only", device, prop.major,
                         prop.minor);
    oip_ctx *c = new (std::nothrow) oip_ctx();
    if (!c) return oip::fail(OIP_E_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (!own_stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return oip::fail(OIP_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    e = cudaMalloc(&c->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
    if (e != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return oip::fail(OIP_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    *out = c;
    return OIP_OK;
}

void oip_ctx_destroy(oip_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    oip::host_pipe_destroy(ctx);
    oip::stt::destroy(ctx);
    oip::downlink_destroy(ctx);
    for (oip_pan_plan &pl : ctx->pan_plans) {
        if (pl.d_plan) cudaFree(pl.d_plan);
        if (pl.h_stage) cudaFreeHost(pl.h_stage);
        if (pl.done) cudaEventDestroy(pl.done);
    }
    if (ctx->d_mss_plan) cudaFree(ctx->d_mss_plan);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->d_err) cudaFree(ctx->d_err);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int oip_ctx_set_option(oip_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return oip::fail(OIP_E_INVALID, "oip_ctx_set_option: null argument");
    if (!strcmp(name, "pan_fast")) ctx->pan_fast = value != 0;
    else if (!strcmp(name, "pan_fast_stages") && value >= 2 && value <= 8) ctx->pan_fast_stages = (int)value;
    else if (!strcmp(name, "pan_fast_minb") && value >= 3 && value <= 4) ctx->pan_fast_minb = (int)value; // kept for old scripts: one variant is built
    else if (!strcmp(name, "host_block_rows") && value >= 64 && value <= (1 << 20)) ctx->host_block_rows = (int)value;
    else if (!strcmp(name, "mss_fast")) ctx->mss_fast = value != 0;
    else if (!strcmp(name, "aos_fused")) ctx->aos_fused = value != 0;
    else if (!strcmp(name, "imtr_runs")) ctx->imtr_runs = value != 0;
    else if (!strcmp(name, "downlink_threads")) ctx->downlink_threads = value != 0;
    else if (!strcmp(name, "mss_fast_rows") && value >= 16 && value <= 32768) ctx->mss_fast_rows = (int)value;
    else if (!strcmp(name, "pan_fast_rows") && value >= 16 && value <= 32768) ctx->pan_fast_rows = (int)value;
    else return oip::fail(OIP_E_INVALID, "unknown option or value out of range: %s=%lld", name, (long long)value);
    for (oip_pan_plan &pl : ctx->pan_plans) pl.key.clear(); // buffers are reused
    ctx->mss_plan_key.clear();
    return OIP_OK;
}

int oip_ctx_sync(oip_ctx *ctx)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    return OIP_OK;
}

void *oip_ctx_stream(oip_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t oip_ctx_launch_count(oip_ctx *ctx) { return ctx ? ctx->launches : 0; }

int oip_dev_alloc(oip_ctx *ctx, size_t bytes, void **d_ptr)
{
    OIP_CHECK_CTX(ctx);
    if (!d_ptr) return oip::fail(OIP_E_INVALID, "null out pointer");
    OIP_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return OIP_OK;
}
int oip_dev_free(oip_ctx *ctx, void *d_ptr)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    OIP_CUDA(cudaFree(d_ptr));
    return OIP_OK;
}
int oip_host_alloc_pinned(size_t bytes, void **h_ptr)
{
    if (!h_ptr) return oip::fail(OIP_E_INVALID, "null out pointer");
    OIP_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
    return OIP_OK;
}
int oip_host_free_pinned(void *h_ptr)
{
    OIP_CUDA(cudaFreeHost(h_ptr));
    return OIP_OK;
}
int oip_copy_h2d(oip_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return OIP_OK;
}
int oip_copy_d2h(oip_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return OIP_OK;
}
int oip_memset_d(oip_ctx *ctx, void *d_dst, int value, size_t bytes)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return OIP_OK;
}

int oip_ipc_export(oip_ctx *ctx, void *d_ptr, uint8_t handle[64])
{
    OIP_CHECK_CTX(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    OIP_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return OIP_OK;
}
int oip_ipc_open(oip_ctx *ctx, const uint8_t handle[64], void **d_peer_ptr)
{
    OIP_CHECK_CTX(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    OIP_CUDA(cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OIP_OK;
}
int oip_ipc_close(oip_ctx *ctx, void *d_peer_ptr)
{
    OIP_CHECK_CTX(ctx);
    OIP_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
    return OIP_OK;
}

/* ref imageop.h:140-192 (host-side text parsing; same fgets/atoi/sscanf sequence) */
int oip_load_rrc_csv(const char *path, int expected, double *kb)
{
    if (!path || !kb) return oip::fail(OIP_E_INVALID, "oip_load_rrc_csv: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return oip::fail(OIP_E_IO, "open RRC Param file failed: %s", path);
    char buff[1024];
    auto bail = [&](int code, const char *msg) {
        fclose(f);
        return oip::fail(code, "%s (%s)", msg, path);
    };
    if (!fgets(buff, sizeof buff, f)) return bail(OIP_E_IO, "LoadRRCParamFile([1]): read file content failed");
    if (!fgets(buff, sizeof buff, f)) return bail(OIP_E_IO, "LoadRRCParamFile([2]): read file content failed");
    int lines = atoi(buff);
    if (lines != expected) {
        fclose(f);
        return oip::fail(OIP_E_INVALID, "LoadRRCParamFile([2]): expected %d lines while %d found in file content",
                         expected, lines);
    }
    if (!fgets(buff, sizeof buff, f)) return bail(OIP_E_IO, "LoadRRCParamFile([3]): read file content failed");
    int index = 0;
    double k = 0, b = 0;
    for (; fgets(buff, sizeof buff, f); ++index) {
        if (sscanf(buff, " %lf , %lf", &k, &b) != 2) {
            fclose(f);
            return oip::fail(OIP_E_INVALID, "line #%d of RRC param file [%s] found invalid", index, path);
        }
        if (index < expected) {
            kb[2 * index] = k;
            kb[2 * index + 1] = b;
        }
    }
    fclose(f);
    if (index != expected)
        return oip::fail(OIP_E_INVALID, "RRC Param file [%s] invalid: %d lines of param expected, %d lines parsed.", path,
                         expected, index);
    return OIP_OK;
}

} // extern "C"
