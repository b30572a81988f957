// host_pipeline.cu -- oip_pan_pipeline_host: the fused PAN path for HOST buffers.
// The strip is cut into row blocks; block b+1's source rows (plus the few halo / stale rows the
// shift needs) travel host->device on a copy stream while block b runs on the compute stream and
// block b-1's output travels device->host on a third stream (pinned buffers make all three
// overlap).  This is what the CLI uses for files and what bench.py reports as "e2e".
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "oip_common.cuh"

namespace oip {

constexpr int HP_SLOTS = 3; // 2, 4 and 6 slots were measured: no difference (the link is the limit)

struct HostPipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    struct Slot {
        void *d_in[8] = {};
        size_t in_cap[8] = {};
        void *d_out = nullptr;
        size_t out_cap = 0;
        cudaEvent_t in_ready = nullptr, compute_done = nullptr, out_done = nullptr;
    } slot[HP_SLOTS];
    void *d_kb[8] = {};
    size_t kb_cap[8] = {};
};

static int hp_get(oip_ctx *ctx, HostPipe **out)
{
    if (!ctx->host_pipe) {
        HostPipe *hp = new (std::nothrow) HostPipe();
        if (!hp) return fail(OIP_E_NOMEM, "out of host memory");
        OIP_CUDA(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
        OIP_CUDA(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
        for (auto &s : hp->slot) {
            OIP_CUDA(cudaEventCreateWithFlags(&s.in_ready, cudaEventDisableTiming));
            OIP_CUDA(cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
            OIP_CUDA(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
        }
        ctx->host_pipe = hp;
    }
    *out = (HostPipe *)ctx->host_pipe;
    return OIP_OK;
}

static int hp_reserve(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return OIP_OK;
    if (*p) OIP_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    OIP_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    return OIP_OK;
}

void host_pipe_destroy(oip_ctx *ctx)
{
    HostPipe *hp = (HostPipe *)ctx->host_pipe;
    if (!hp) return;
    for (auto &s : hp->slot) {
        for (int i = 0; i < 8; ++i)
            if (s.d_in[i]) cudaFree(s.d_in[i]);
        if (s.d_out) cudaFree(s.d_out);
        if (s.in_ready) cudaEventDestroy(s.in_ready);
        if (s.compute_done) cudaEventDestroy(s.compute_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    for (int i = 0; i < 8; ++i)
        if (hp->d_kb[i]) cudaFree(hp->d_kb[i]);
    if (hp->h2d) cudaStreamDestroy(hp->h2d);
    if (hp->d2h) cudaStreamDestroy(hp->d2h);
    delete hp;
    ctx->host_pipe = nullptr;
}

// rows x row_bytes between pitched buffers; one linear copy when both sides are dense (the 2-D DMA path is slower:
// 34 GB/s against 47 GB/s per direction on this box, tools/pcie_probe.py)
static cudaError_t copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows,
                             cudaMemcpyKind kind, cudaStream_t st)
{
    if (dpitch == row_bytes && spitch == row_bytes) return cudaMemcpyAsync(dst, src, row_bytes * rows, kind, st);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, kind, st);
}

static int64_t row_bytes(int fmt, int w)
{
    switch (fmt) {
    case OIP_FMT_PACK12: return ((int64_t)w * 12 + 7) / 8;
    case OIP_FMT_PACK10: return ((int64_t)w * 10 + 7) / 8;
    default: return (int64_t)w * 2;
    }
}

} // namespace oip

using namespace oip;

extern "C" int oip_pan_pipeline_host(oip_ctx *ctx, const oip_pan_desc *h)
{
    OIP_CHECK_CTX(ctx);
    if (!h) return fail(OIP_E_INVALID, "null descriptor");
    if (h->n_ccd < 1 || h->n_ccd > 8) return fail(OIP_E_INVALID, "n_ccd=%d out of range 1..8", h->n_ccd);
    if (!h->d_out && h->n_rows > 0) return fail(OIP_E_INVALID, "output buffer is null");
    for (int i = 0; i < h->n_ccd; ++i) {
        if (h->ccd[i].fmt == OIP_FMT_BE16_TILES) return fail(OIP_E_UNSUPPORTED, "host pipeline takes line formats only");
        if (h->ccd[i].n_seg < 1 || h->ccd[i].n_seg > OIP_MAX_SEG) return fail(OIP_E_INVALID, "ccd %d: n_seg=%d", i, h->ccd[i].n_seg);
        for (int s = 0; s < h->ccd[i].n_seg; ++s)
            if (!h->ccd[i].seg[s].base) return fail(OIP_E_INVALID, "ccd %d: host segment %d is null", i, s);
    }
    if (h->n_rows <= 0) return OIP_OK;
    HostPipe *hp = nullptr;
    int rc = hp_get(ctx, &hp);
    if (rc) return rc;
    const int out_w = oip_pan_out_width(h->n_ccd, h->w, h->fold_half);

    // RRC coefficients: tiny, once per call, on the compute stream
    const double *d_kb[8] = {};
    for (int i = 0; i < h->n_ccd; ++i) {
        if (!h->ccd[i].d_kb) continue;
        const size_t bytes = (size_t)h->w * 16;
        rc = hp_reserve(&hp->d_kb[i], &hp->kb_cap[i], bytes);
        if (rc) return rc;
        OIP_CUDA(cudaMemcpyAsync(hp->d_kb[i], h->ccd[i].d_kb, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_kb[i] = (const double *)hp->d_kb[i];
    }

    // block schedule: small blocks at both ends (the first H2D and the last D2H cannot overlap with anything), the
    // configured size in between
    const int64_t HP_BLOCK_ROWS = ctx->host_block_rows;
    std::vector<int64_t> sizes;
    {
        int64_t left = h->n_rows;
        std::vector<int64_t> head, tail;
        for (int64_t b = std::min<int64_t>(256, HP_BLOCK_ROWS); b < HP_BLOCK_ROWS && left >= 4 * b; b *= 2) {
            head.push_back(b);
            tail.push_back(b);
            left -= 2 * b;
        }
        sizes = head;
        while (left > 0) { sizes.push_back(std::min(left, HP_BLOCK_ROWS)); left -= sizes.back(); }
        sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
    }
    int blk = 0;
    int64_t r0 = h->row0;
    for (size_t bi = 0; bi < sizes.size(); r0 += sizes[bi], ++bi, ++blk) {
        const int64_t nr = sizes[bi];
        HostPipe::Slot &S = hp->slot[blk % HP_SLOTS];
        oip_pan_desc d = *h;
        d.row0 = r0;
        d.n_rows = nr;
        // ---- H2D of the rows this block reads (own rows + halo + stale-section rows)
        OIP_CUDA(cudaStreamWaitEvent(hp->h2d, S.compute_done, 0)); // the slot's previous kernel is done with d_in
        for (int i = 0; i < h->n_ccd; ++i) {
            const oip_ccd_src &src = h->ccd[i];
            int64_t rg[2 * OIP_MAX_SEG];
            int n_rg = 0;
            rc = oip_pan_row_ranges(&d, i, rg, OIP_MAX_SEG, &n_rg);
            if (rc) return rc;
            // a scanline-block shard hands in up to OIP_MAX_SEG HOST segments (its own block with the halo rows read from
            // the file, the stale rows of a partial last section): every needed range is cut against each of them
            const int64_t rb = row_bytes(src.fmt, h->w);
            const int64_t pitch_d = (rb + 15) / 16 * 16;
            struct Piece { int64_t a, n; int seg; };
            Piece pieces[OIP_MAX_SEG * OIP_MAX_SEG];
            int n_pc = 0;
            int64_t n_all = 0;
            for (int k = 0; k < n_rg; ++k)
                for (int q = 0; q < src.n_seg; ++q) {
                    const oip_row_seg &hs = src.seg[q];
                    const int64_t a = std::max(rg[2 * k], hs.row0), b = std::min(rg[2 * k + 1], hs.row0 + hs.n_rows);
                    if (b <= a) continue;
                    bool dup = false; // host segments may overlap (halo rows of a block are also inside the stale range)
                    for (int e = 0; e < n_pc; ++e) dup = dup || (a >= pieces[e].a && b <= pieces[e].a + pieces[e].n);
                    if (dup) continue;
                    pieces[n_pc++] = {a, b - a, q};
                    n_all += b - a;
                }
            if (n_pc > OIP_MAX_SEG) return fail(OIP_E_INVALID, "ccd %d: a row block needs %d host pieces (max %d)", i, n_pc, OIP_MAX_SEG);
            rc = hp_reserve(&S.d_in[i], &S.in_cap[i], (size_t)(std::max<int64_t>(HP_BLOCK_ROWS + 256, n_all) * pitch_d));
            if (rc) return rc;
            uint8_t *dst = (uint8_t *)S.d_in[i];
            oip_ccd_src &o = d.ccd[i];
            o.d_kb = d_kb[i];
            o.n_seg = 0;
            for (int e = 0; e < n_pc; ++e) {
                const oip_row_seg &hs = src.seg[pieces[e].seg];
                const int64_t a = pieces[e].a, n = pieces[e].n;
                OIP_CUDA(copy_rows(dst, (size_t)pitch_d, (const uint8_t *)hs.base + (a - hs.row0) * hs.pitch_bytes, (size_t)hs.pitch_bytes,
                                   (size_t)rb, (size_t)n, cudaMemcpyHostToDevice, hp->h2d));
                o.seg[o.n_seg++] = {dst, a, n, pitch_d};
                dst += n * pitch_d;
            }
            if (o.n_seg == 0) { // nothing to read (can only happen for degenerate geometry): keep a valid segment
                o.seg[0] = {S.d_in[i], 0, 0, pitch_d};
                o.n_seg = 1;
            }
        }
        OIP_CUDA(cudaEventRecord(S.in_ready, hp->h2d));
        // ---- compute
        rc = hp_reserve(&S.d_out, &S.out_cap, (size_t)(HP_BLOCK_ROWS * out_w * 2));
        if (rc) return rc;
        OIP_CUDA(cudaStreamWaitEvent(ctx->stream, S.in_ready, 0));
        OIP_CUDA(cudaStreamWaitEvent(ctx->stream, S.out_done, 0)); // previous D2H of this slot finished
        d.d_out = (uint16_t *)S.d_out;
        d.out_pitch_px = out_w;
        static const bool skip_kernels = getenv("OIP_HP_SKIP_KERNELS") != nullptr; // timing experiment: copies only
        rc = skip_kernels ? OIP_OK : oip_pan_pipeline(ctx, &d);
        if (rc) return rc;
        OIP_CUDA(cudaEventRecord(S.compute_done, ctx->stream));
        // ---- D2H
        OIP_CUDA(cudaStreamWaitEvent(hp->d2h, S.compute_done, 0));
        OIP_CUDA(copy_rows((uint8_t *)h->d_out + (r0 - h->row0) * h->out_pitch_px * 2, (size_t)h->out_pitch_px * 2, S.d_out,
                           (size_t)out_w * 2, (size_t)out_w * 2, (size_t)nr, cudaMemcpyDeviceToHost, hp->d2h));
        OIP_CUDA(cudaEventRecord(S.out_done, hp->d2h));
    }
    OIP_CUDA(cudaStreamSynchronize(hp->d2h));
    OIP_CUDA(cudaStreamSynchronize(ctx->stream));
    return oip_pan_check_error(ctx);
}
