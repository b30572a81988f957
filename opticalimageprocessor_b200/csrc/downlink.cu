// downlink.cu -- oip_downlink_to_stitched: raw downlink files -> stitched PAN raster in one call ("K9").
//
// The reference runs this as three programs that talk through files: `auxsep` writes .IMDT, .AUX, .PAN.RAW, .MSS.RAW
// (ref aux_separator.h:224-245), `prestitch` writes .RRC.RAW and .PRESTT.RAW (ref stitcher.h:83-146), `stitch` writes the
// product (ref imageop.h:277-363).  Here the chain is: AOS scan -> IMTR re-framing -> image-frame index on each CCD's
// downlink, then ONE fused PAN launch whose fast kernel gathers its rows straight from the sub-images of the IMDT streams
// (OIP_FMT_BE16_TILES): no PAN raster, no corrected raster, no shifted raster is ever written.  What still passes
// through HBM between the kernels is the IMDT stream itself (the CRC-checked 866-byte bodies) and the small tables.
#include <algorithm>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "oip_common.cuh"

namespace oip {

struct DownlinkState {
    struct PerCcd {
        void *d_payload_off = nullptr; size_t off_cap = 0;
        void *d_imdt = nullptr;        size_t imdt_cap = 0;
        void *d_tab = nullptr;         size_t tab_cap = 0;
        int64_t imdt_bytes = 0;
        std::vector<int64_t> h_tab;
        std::vector<oip_frame_entry> entries;
    } c[8];
    // stage 1 of the CCDs runs concurrently: one child context (own stream, own scratch) and one host thread per CCD, so
    // that the small table kernels and the host round trips of one downlink hide behind the streaming kernels of another
    oip_ctx *sub[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in = nullptr;
};

static int dl_reserve(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return OIP_OK;
    if (*p) OIP_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    OIP_CUDA(cudaMalloc(p, bytes + bytes / 16 + 4096));
    *cap = bytes + bytes / 16 + 4096;
    return OIP_OK;
}

void downlink_destroy(oip_ctx *ctx)
{
    DownlinkState *st = (DownlinkState *)ctx->downlink_state;
    if (!st) return;
    for (auto &c : st->c) {
        if (c.d_payload_off) cudaFree(c.d_payload_off);
        if (c.d_imdt) cudaFree(c.d_imdt);
        if (c.d_tab) cudaFree(c.d_tab);
    }
    for (oip_ctx *sc : st->sub)
        if (sc) oip_ctx_destroy(sc);
    if (st->ev_in) cudaEventDestroy(st->ev_in);
    delete st;
    ctx->downlink_state = nullptr;
}

// stage 1 of one CCD on context cx (the caller's or a child): every byte of the file is searched and CRC-checked exactly
// like `auxsep` does
static int downlink_stage1(oip_ctx *cx, const oip_downlink_desc *d, int i, DownlinkState::PerCcd &c, oip_downlink_stats &ls)
{
    const oip_downlink_src &s = d->ccd[i];
    ls = oip_downlink_stats{};
    const size_t cap_off = s.n_bytes / 1024 + 1;
    int rc = dl_reserve(&c.d_payload_off, &c.off_cap, cap_off * 8);
    if (rc) return rc;
    rc = oip_aos_scan(cx, s.d_file, s.n_bytes, (uint64_t *)c.d_payload_off, cap_off, ls.aos);          // ref aux_separator.h:395-467
    if (rc) return rc;
    const int64_t n_valid = ls.aos[0];
    const size_t cap_imdt = (size_t)(n_valid * 880 / 882 + 1) * 866 + 64;
    rc = dl_reserve(&c.d_imdt, &c.imdt_cap, cap_imdt);
    if (rc) return rc;
    rc = oip_imtr_deframe(cx, s.d_file, (const uint64_t *)c.d_payload_off, n_valid, (uint8_t *)c.d_imdt, cap_imdt, ls.imtr,
                          &ls.imdt_bytes);                                                              // :469-590
    if (rc) return rc;
    // frame index: one call with a capacity that holds every complete frame the stream can contain plus the
    // zero-filled gap frames of a 16-bit sequence counter; grown once if a pathological stream needs more
    const int64_t frame_bytes = 192ll * d->geom.tile_lines + 40ll * d->geom.tile_lines * d->geom.tile_cols * 2 + 172;
    int64_t cap_fr = ls.imdt_bytes / frame_bytes + 64;
    for (int attempt = 0; attempt < 2; ++attempt) {
        c.entries.resize((size_t)cap_fr);
        rc = oip_image_frames_index(cx, (const uint8_t *)c.d_imdt, (size_t)ls.imdt_bytes, &d->geom, c.entries.data(), cap_fr, ls.frames); // :627-656, :287-320
        if (rc == OIP_E_INVALID && ls.frames[1] > cap_fr) { cap_fr = ls.frames[1]; continue; }
        break;
    }
    if (rc) return rc;
    c.imdt_bytes = ls.imdt_bytes;
    return OIP_OK;
}

} // namespace oip

using namespace oip;

extern "C" int oip_downlink_to_stitched(oip_ctx *ctx, const oip_downlink_desc *d, int64_t *rows_out, oip_downlink_stats *stats)
{
    OIP_CHECK_CTX(ctx);
    if (rows_out) *rows_out = 0;
    if (!d) return fail(OIP_E_INVALID, "oip_downlink_to_stitched: null descriptor");
    if (d->n_ccd < 1 || d->n_ccd > 8) return fail(OIP_E_INVALID, "n_ccd=%d out of range 1..8", d->n_ccd);
    if (d->geom.tile_cols < 1 || d->geom.tile_lines < 1) return fail(OIP_E_INVALID, "oip_downlink_to_stitched: bad frame geometry");
    if (!d->d_out) return fail(OIP_E_INVALID, "oip_downlink_to_stitched: d_out is null");
    DownlinkState *st = (DownlinkState *)ctx->downlink_state;
    if (!st) ctx->downlink_state = st = new (std::nothrow) DownlinkState();
    if (!st) return fail(OIP_E_NOMEM, "out of host memory");
    const int w = 8 * d->geom.tile_cols, lpf = 4 * d->geom.tile_lines;
    int64_t n_frames = -1;
    int rc;
    for (int i = 0; i < d->n_ccd; ++i)
        if (!d->ccd[i].d_file || d->ccd[i].n_bytes < 1024) return fail(OIP_E_INVALID, "ccd %d: empty downlink", i);
    // ---- stage 1 per CCD, the CCDs side by side (option downlink_threads = 0: one after the other on the caller's stream)
    oip_downlink_stats ls[8];
    const bool side_by_side = ctx->downlink_threads && d->n_ccd > 1;
    if (side_by_side) {
        if (!st->ev_in) OIP_CUDA(cudaEventCreateWithFlags(&st->ev_in, cudaEventDisableTiming));
        OIP_CUDA(cudaEventRecord(st->ev_in, ctx->stream));    // the files were produced on the caller's stream
        for (int i = 0; i < d->n_ccd; ++i) {
            if (!st->sub[i]) {
                rc = oip_ctx_create(ctx->device, nullptr, 1, &st->sub[i]);
                if (rc) return rc;
            }
            st->sub[i]->aos_fused = ctx->aos_fused;
            st->sub[i]->imtr_runs = ctx->imtr_runs;
            OIP_CUDA(cudaStreamWaitEvent(st->sub[i]->stream, st->ev_in, 0));
        }
        int rcs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::string msgs[8];
        auto work = [&](int i) {
            if (cudaSetDevice(ctx->device) != cudaSuccess) { rcs[i] = OIP_E_CUDA; msgs[i] = "cudaSetDevice failed"; return; }
            rcs[i] = downlink_stage1(st->sub[i], d, i, st->c[i], ls[i]);
            if (rcs[i]) msgs[i] = oip_last_error();           // the error text is per thread
        };
        std::vector<std::thread> th;
        int started = 1;                                      // CCD 0 runs on the calling thread
        try {
            for (int i = 1; i < d->n_ccd; ++i) { th.emplace_back(work, i); started = i + 1; }
        } catch (const std::exception &) {                    // no more threads to be had: the rest runs here, one after the other
        }
        work(0);
        for (int i = started; i < d->n_ccd; ++i) work(i);
        for (std::thread &t : th) t.join();
        for (int i = 0; i < d->n_ccd; ++i) {
            ctx->launches += st->sub[i]->launches;
            st->sub[i]->launches = 0;
        }
        for (int i = 0; i < d->n_ccd; ++i)
            if (rcs[i]) return fail(rcs[i], "ccd %d: %s", i, msgs[i].c_str());
        // every stage-1 call ended with a synchronisation of its stream: the IMDT streams are complete for what follows
    } else {
        for (int i = 0; i < d->n_ccd; ++i) {
            rc = downlink_stage1(ctx, d, i, st->c[i], ls[i]);
            if (rc) return rc;
        }
    }
    for (int i = 0; i < d->n_ccd; ++i) {
        if (stats) stats[i] = ls[i];
        n_frames = n_frames < 0 ? ls[i].frames[1] : std::min<int64_t>(n_frames, ls[i].frames[1]);
    }
    if (n_frames <= 0) return OIP_OK;
    const int64_t rows = n_frames * lpf;
    if (d->out_rows_cap < rows) return fail(OIP_E_INVALID, "oip_downlink_to_stitched: %lld lines exceed the output capacity %lld", (long long)rows, (long long)d->out_rows_cap);
    // ---- stages 2 + 3: one fused PAN launch over the frame tiles
    oip_pan_desc p{};
    p.n_ccd = d->n_ccd; p.w = w; p.total_rows = rows; p.row0 = 0; p.n_rows = rows;
    p.fold_half = d->fold_half; p.section_rows = d->section_rows; p.row_guard = d->row_guard;
    p.d_out = d->d_out; p.out_pitch_px = d->out_pitch_px;
    for (int i = 0; i < d->n_ccd; ++i) {
        DownlinkState::PerCcd &c = st->c[i];
        c.h_tab.resize((size_t)n_frames * 40);
        for (int64_t f = 0; f < n_frames; ++f)
            for (int k = 0; k < 40; ++k) c.h_tab[(size_t)f * 40 + k] = c.entries[(size_t)f].tile_off[k];
        rc = dl_reserve(&c.d_tab, &c.tab_cap, c.h_tab.size() * 8);
        if (rc) return rc;
        OIP_CUDA(cudaMemcpyAsync(c.d_tab, c.h_tab.data(), c.h_tab.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        oip_ccd_src &o = p.ccd[i];
        o.fmt = OIP_FMT_BE16_TILES; o.n_seg = 1;
        o.seg[0] = {c.d_imdt, 0, rows, 0};
        o.d_kb = d->ccd[i].d_kb; o.shifted = d->ccd[i].shifted; o.dX = d->ccd[i].dX; o.dY = d->ccd[i].dY;
        o.d_tile_off = (const int64_t *)c.d_tab; o.h_tile_off = c.h_tab.data();
        o.tile_cols = d->geom.tile_cols; o.tile_lines = d->geom.tile_lines;
    }
    rc = oip_pan_pipeline(ctx, &p);
    if (rc) return rc;
    // ---- the other products of `auxsep`, on request: aux blocks and MSS lines (ref aux_separator.h:335-339, :341-364)
    for (int i = 0; i < d->n_ccd; ++i) {
        if (!d->d_aux[i] && !d->d_mss[i]) continue;
        DownlinkState::PerCcd &c = st->c[i];
        rc = oip_unpack_frames(ctx, (const uint8_t *)c.d_imdt, (size_t)c.imdt_bytes, &d->geom, c.entries.data(), n_frames, d->d_aux[i], nullptr, d->d_mss[i]);
        if (rc) return rc;
    }
    if (rows_out) *rows_out = rows;
    return OIP_OK;
}
