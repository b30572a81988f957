// pixel_ops.cu -- stand-alone per-pixel kernels: in-place RRC, line unpacking, 4-channel concat.
#include "oip_common.cuh"

namespace oip {

// ---------------------------------------------------------------------------------------------
// InplaceRRC (ref imageop.h:129-138).  Column-persistent threads: a thread owns 8 adjacent
// detectors, keeps their (k,b) in registers and walks down the rows with 128-bit loads/stores.
// HBM-bound: 2 B read + 2 B written per pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rrc_vec8_kernel(uint16_t *__restrict__ img, int w8, int64_t h, int64_t pitch,
                                                       const double2 *__restrict__ kb, int64_t rows_per_block)
{
    const int cg = blockIdx.x * blockDim.x + threadIdx.x;
    if (cg >= w8) return;
    double k[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double2 v = kb[cg * 8 + i];
        k[i] = v.x;
        b[i] = v.y;
    }
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(h, r0 + rows_per_block);
    uint16_t *p = img + r0 * pitch + (int64_t)cg * 8;
    int64_t r = r0;
    // (four rows in flight per thread were measured: 1205 against 1252 Gpixel/s for two)
    for (; r + 1 < r1; r += 2, p += 2 * pitch) {
        uint4 a = ldg_nc_v4(p), c = ldg_nc_v4(p + pitch);
        uint32_t *aw = reinterpret_cast<uint32_t *>(&a), *cw = reinterpret_cast<uint32_t *>(&c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            aw[i] = rrc_px(aw[i] & 0xFFFFu, k[2 * i], b[2 * i]) | (rrc_px(aw[i] >> 16, k[2 * i + 1], b[2 * i + 1]) << 16);
            cw[i] = rrc_px(cw[i] & 0xFFFFu, k[2 * i], b[2 * i]) | (rrc_px(cw[i] >> 16, k[2 * i + 1], b[2 * i + 1]) << 16);
        }
        stg_na_v4(p, a);
        stg_na_v4(p + pitch, c);
    }
    if (r < r1) {
        uint4 a = ldg_nc_v4(p);
        uint32_t *aw = reinterpret_cast<uint32_t *>(&a);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            aw[i] = rrc_px(aw[i] & 0xFFFFu, k[2 * i], b[2 * i]) | (rrc_px(aw[i] >> 16, k[2 * i + 1], b[2 * i + 1]) << 16);
        stg_na_v4(p, a);
    }
}

__global__ void rrc_scalar_kernel(uint16_t *img, int w, int64_t h, int64_t pitch, const double2 *kb, int x0)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const double2 v = kb[x];
    for (int64_t r = blockIdx.y; r < h; r += gridDim.y) {
        uint16_t *p = img + r * pitch + x;
        *p = (uint16_t)rrc_px(*p, v.x, v.y);
    }
}

// ---------------------------------------------------------------------------------------------
// unpack lines: BE16 swap / MSB-first packed 12- and 10-bit -> u16 LE  (extension)
// one thread = 8 output pixels (16 B store); input 16 / 12 / 10 bytes
// ---------------------------------------------------------------------------------------------
__global__ void unpack_lines_kernel(const uint8_t *__restrict__ in, int fmt, int w, int64_t rows, int64_t pitch,
                                    uint16_t *__restrict__ out)
{
    const int64_t n8 = (int64_t)(w + 7) / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8 * rows; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n8;
        const int x = (int)(i - r * n8) * 8;
        const uint8_t *row = in + r * pitch;
        uint16_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = x + j;
            uint32_t s = 0;
            if (c < w) {
                if (fmt == OIP_FMT_BE16) s = ((uint32_t)row[2 * c] << 8) | row[2 * c + 1];
                else if (fmt == OIP_FMT_LE16) s = ((uint32_t)row[2 * c + 1] << 8) | row[2 * c];
                else if (fmt == OIP_FMT_PACK12) {
                    const uint8_t *p = row + (int64_t)(c >> 1) * 3;
                    s = (c & 1) ? (((uint32_t)(p[1] & 0x0F) << 8) | p[2]) : (((uint32_t)p[0] << 4) | (p[1] >> 4));
                } else {
                    const uint8_t *p = row + (int64_t)(c >> 2) * 5;
                    int k = c & 3;
                    s = (((uint32_t)p[k] << (2 + 2 * k)) | ((uint32_t)p[k + 1] >> (6 - 2 * k))) & 0x3FF;
                }
            }
            v[j] = (uint16_t)s;
        }
        uint16_t *o = out + r * (int64_t)w + x;
        if (x + 8 <= w && (((uintptr_t)o) & 15) == 0) {
            *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<uint4 *>(v);
        } else {
            for (int j = 0; j < 8 && x + j < w; ++j) o[j] = v[j];
        }
    }
}

// word path for the packed formats: one thread = 16 output pixels = 6 (12-bit) or 5 (10-bit) aligned input words.
// The MSB-first bit stream is rebuilt as big-endian words (one PRMT each); pixel k sits BITS*k bits from the top of the
// group: one funnel shift + one right shift per pixel, two 16-byte stores per thread.
template <int BITS>
__global__ void __launch_bounds__(256) unpack_packed16_kernel(const uint8_t *__restrict__ in, int w, int64_t rows, int64_t pitch,
                                                              uint16_t *__restrict__ out)
{
    constexpr int NW = BITS * 16 / 32;
    const int64_t n16 = (int64_t)w / 16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16 * rows; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n16;
        const int g = (int)(i - r * n16);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in + r * pitch) + (int64_t)g * NW;
        uint32_t be[NW + 1];
#pragma unroll
        for (int k = 0; k < NW; ++k) be[k] = __byte_perm(__ldg(src + k), 0u, 0x0123);
        be[NW] = 0u;
        uint32_t px[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int bit = BITS * k;
            px[k] = __funnelshift_l(be[(bit >> 5) + 1], be[bit >> 5], bit & 31) >> (32 - BITS);
        }
        uint4 *o = reinterpret_cast<uint4 *>(out + r * (int64_t)w + 16 * (int64_t)g);
        stg_na_v4(o, make_uint4(px[0] | (px[1] << 16), px[2] | (px[3] << 16), px[4] | (px[5] << 16), px[6] | (px[7] << 16)));
        stg_na_v4(o + 1, make_uint4(px[8] | (px[9] << 16), px[10] | (px[11] << 16), px[12] | (px[13] << 16), px[14] | (px[15] << 16)));
    }
}

// ---------------------------------------------------------------------------------------------
// StitchTiff geometry on CV_16UC4 pixels with band map (ref imageop.h:416-421, :501-506, :529)
// one thread = one 4-channel pixel (8 B)
// ---------------------------------------------------------------------------------------------
struct ConcatC4 {
    const uint16_t *img[8];
    int n, w, f;
    int64_t rows;
    int map[4];
};
__global__ void __launch_bounds__(256) concat_c4_kernel(const __grid_constant__ ConcatC4 P, uint16_t *__restrict__ out)
{
    // grid: x over the output pixels of a line, y (strided) over the lines; the channel map as two PRMT selectors
    const int wout = P.n * P.w - 2 * (P.n - 1) * P.f;
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    if (xo >= wout) return;
    int k = 0, x = xo; // which image: widths are w-f, w-2f ..., w-f
    const int first = P.n == 1 ? P.w : P.w - P.f;
    if (x >= first) {
        x -= first;
        const int mid = P.w - 2 * P.f;
        k = 1 + (mid > 0 ? x / mid : 0);
        if (k > P.n - 1) k = P.n - 1;
        x -= (k - 1) * mid;
        x += P.f;
    }
    uint32_t sel[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m0 = P.map[2 * h], m1 = P.map[2 * h + 1]; // output halfwords 2h, 2h+1 <- input halfwords m0, m1
        sel[h] = (uint32_t)(2 * m0) | ((uint32_t)(2 * m0 + 1) << 4) | ((uint32_t)(2 * m1) << 8) | ((uint32_t)(2 * m1 + 1) << 12);
    }
    const uint16_t *src = P.img[k] + (int64_t)x * 4;
    const int64_t in_pitch = (int64_t)P.w * 4, out_pitch = (int64_t)wout * 4;
    for (int64_t y = blockIdx.y; y < P.rows; y += gridDim.y) {
        uint2 s;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(s.x), "=r"(s.y) : "l"(src + y * in_pitch));
        uint2 o;
        o.x = __byte_perm(s.x, s.y, sel[0]);
        o.y = __byte_perm(s.x, s.y, sel[1]);
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out + y * out_pitch + (int64_t)xo * 4), "r"(o.x), "r"(o.y) : "memory");
    }
}

} // namespace oip

using namespace oip;

extern "C" int oip_rrc_u16(oip_ctx *ctx, uint16_t *d_img, int w, int64_t h, int64_t pitch_px, const double *d_kb)
{
    OIP_CHECK_CTX(ctx);
    if (!d_img || !d_kb) return fail(OIP_E_INVALID, "oip_rrc_u16: null pointer");
    if (w <= 0 || h < 0 || pitch_px < w) return fail(OIP_E_INVALID, "oip_rrc_u16: bad geometry w=%d h=%lld pitch=%lld", w, (long long)h, (long long)pitch_px);
    if (h == 0) return OIP_OK;
    const bool vec = (((uintptr_t)d_img & 15) == 0) && (pitch_px % 8 == 0) && (((uintptr_t)d_kb & 15) == 0);
    const int w8 = vec ? w / 8 : 0;
    if (w8 > 0) {
        const int bx = (w8 + 127) / 128;
        int by = (int)std::min<int64_t>(h, std::max<int64_t>(1, (int64_t)ctx->sm_count * 16 / bx));
        int64_t rpb = (h + by - 1) / by;
        rpb = (rpb + 1) & ~(int64_t)1;
        by = (int)((h + rpb - 1) / rpb);
        rrc_vec8_kernel<<<dim3(bx, by), 128, 0, ctx->stream>>>(d_img, w8, h, pitch_px, (const double2 *)d_kb, rpb);
        OIP_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    if (w8 * 8 < w) {
        const int rem = w - w8 * 8;
        rrc_scalar_kernel<<<dim3((rem + 127) / 128, (unsigned)std::min<int64_t>(h, 2048)), 128, 0, ctx->stream>>>(
            d_img, w, h, pitch_px, (const double2 *)d_kb, w8 * 8);
        OIP_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    return OIP_OK;
}

extern "C" int oip_unpack_lines(oip_ctx *ctx, const void *d_in, int fmt, int w, int64_t rows, int64_t pitch_bytes,
                                uint16_t *d_out)
{
    OIP_CHECK_CTX(ctx);
    if (!d_in || !d_out) return fail(OIP_E_INVALID, "oip_unpack_lines: null pointer");
    if (fmt < OIP_FMT_LE16 || fmt > OIP_FMT_PACK10) return fail(OIP_E_INVALID, "oip_unpack_lines: format %d", fmt);
    if (w <= 0 || rows < 0) return fail(OIP_E_INVALID, "oip_unpack_lines: bad geometry");
    if (rows == 0) return OIP_OK;
    const int64_t n = ((int64_t)w + 7) / 8 * rows;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 32);
    const bool word_path = (fmt == OIP_FMT_PACK12 || fmt == OIP_FMT_PACK10) && w % 16 == 0 && pitch_bytes % 4 == 0 &&
                           (((uintptr_t)d_in) & 3) == 0 && (((uintptr_t)d_out) & 15) == 0;
    if (word_path) {
        const int b2 = (int)std::min<int64_t>(((int64_t)w / 16 * rows + 255) / 256, (int64_t)ctx->sm_count * 32);
        if (fmt == OIP_FMT_PACK12) unpack_packed16_kernel<12><<<b2, 256, 0, ctx->stream>>>((const uint8_t *)d_in, w, rows, pitch_bytes, d_out);
        else unpack_packed16_kernel<10><<<b2, 256, 0, ctx->stream>>>((const uint8_t *)d_in, w, rows, pitch_bytes, d_out);
    } else
        unpack_lines_kernel<<<blocks, 256, 0, ctx->stream>>>((const uint8_t *)d_in, fmt, w, rows, pitch_bytes, d_out);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}

extern "C" int oip_stitch_concat_c4(oip_ctx *ctx, const uint16_t *const *d_img, int n_img, int w, int64_t rows,
                                    int fold_half, const int *band_map, uint16_t *d_dst)
{
    OIP_CHECK_CTX(ctx);
    if (!d_img || !d_dst) return fail(OIP_E_INVALID, "oip_stitch_concat_c4: null pointer");
    if (n_img < 1 || n_img > 8) return fail(OIP_E_INVALID, "n_img=%d out of range 1..8", n_img);
    if (w <= 0 || fold_half < 0 || 2 * fold_half >= w) return fail(OIP_E_INVALID, "bad w/fold");
    ConcatC4 P{};
    for (int i = 0; i < n_img; ++i) {
        if (!d_img[i]) return fail(OIP_E_INVALID, "image %d is null", i);
        P.img[i] = d_img[i];
    }
    P.n = n_img; P.w = w; P.f = fold_half; P.rows = rows;
    for (int b = 0; b < 4; ++b) {
        int m = band_map ? band_map[b] : b + 1;
        if (m < 1 || m > 4) return fail(OIP_E_INVALID, "invalid band index"); /* ref main.cpp:183-187 */
        P.map[b] = m - 1;
    }
    if (rows == 0) return OIP_OK;
    const int wout = oip_pan_out_width(n_img, w, fold_half);
    const int bx = (wout + 255) / 256;
    const int by = (int)std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)ctx->sm_count * 32 / bx));
    concat_c4_kernel<<<dim3(bx, by), 256, 0, ctx->stream>>>(P, d_dst);
    OIP_CUDA(cudaGetLastError());
    ctx->launches++;
    return OIP_OK;
}
