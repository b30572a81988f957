// stt.cu -- SURVEY 8(f) N1: inter-CMOS offset estimation, the caller that produces (dX, dY) for the PAN path.
// Replaces Stitcher::CalcSttParameters (ref stitcher.h:148-201) and the cv::phaseCorrelate it calls (stitcher.h:180;
// OpenCV imgproc/src/phasecorr.cpp): u16 overlap columns -> float, zero-pad to the optimal DFT size, forward DFTs,
// normalised cross-power spectrum, inverse DFT, quadrant swap, first maximum, 5x5 weighted centroid.
// The two DFTs are cuFFT (a plain library transform, like the reference's cv::dft); everything around them is here.
// Floating point: parity with cv2.phaseCorrelate is a tolerance (tests: 2e-3 px), not bits.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cufft.h>

#include "oip_common.cuh"

namespace oip {
namespace stt {

struct State {
    int M = 0, N = 0;
    cufftHandle fwd = 0, inv = 0;
    bool have = false;
};

static int optimal_dft_size(int n) // cv::getOptimalDFTSize: smallest 2^a 3^b 5^c >= n
{
    int64_t best = -1;
    for (int64_t p2 = 1; p2 < 2 * (int64_t)n; p2 *= 2)
        for (int64_t p3 = p2; p3 < 2 * (int64_t)n; p3 *= 3)
            for (int64_t p5 = p3; p5 < 2 * (int64_t)n; p5 *= 5)
                if (p5 >= n && (best < 0 || p5 < best)) best = p5;
    return (int)best;
}

// both slices -> zero-padded float planes [2][M][N]
__global__ void pack_kernel(const uint16_t *__restrict__ a, int64_t pitch_a, const uint16_t *__restrict__ b, int64_t pitch_b,
                            int rows, int cols, int M, int N, float *__restrict__ out)
{
    const int64_t n = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = i >= n;
        const int64_t j = i - (img ? n : 0);
        const int y = (int)(j / N), x = (int)(j - (int64_t)y * N);
        float v = 0.f;
        if (y < rows && x < cols) v = (float)(img ? b[(int64_t)y * pitch_b + x] : a[(int64_t)y * pitch_a + x]);
        out[i] = v;
    }
}

// N2 (ref preproc.h:256-304): plane 0 = the PAN slice as float, plane 1 = the MSS band slice upscaled to the same size by
// cv::resize(..., INTER_CUBIC) on CV_32FC1: f = (float)((d + 0.5) * scale - 0.5), taps floor(f)-1..+2 clamped to the
// image, weights interpolateCubic(A = -0.75) in float, horizontal pass then vertical pass (OpenCV's order)
__device__ __forceinline__ void cubic_w(float x, float (&w)[4])
{
    const float A = -0.75f;
    w[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
    w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    w[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
    w[3] = 1.f - w[0] - w[1] - w[2];
}
__global__ void pack_resize_kernel(const uint16_t *__restrict__ pan, int64_t pitch_pan, const uint16_t *__restrict__ band,
                                   int64_t pitch_band, int rows, int cols, int brows, int bcols, double scale_y, double scale_x,
                                   int M, int N, float *__restrict__ out)
{
    const int64_t n = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = i >= n;
        const int64_t j = i - (img ? n : 0);
        const int y = (int)(j / N), x = (int)(j - (int64_t)y * N);
        float v = 0.f;
        if (y < rows && x < cols) {
            if (!img) {
                v = (float)pan[(int64_t)y * pitch_pan + x];
            } else {
                const float fy = (float)((y + 0.5) * scale_y - 0.5), fx = (float)((x + 0.5) * scale_x - 0.5);
                const int sy = (int)floorf(fy), sx = (int)floorf(fx);
                float wy[4], wx[4];
                cubic_w(fy - (float)sy, wy);
                cubic_w(fx - (float)sx, wx);
                int xi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) xi[k] = min(max(sx - 1 + k, 0), bcols - 1);
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint16_t *row = band + (int64_t)min(max(sy - 1 + r, 0), brows - 1) * pitch_band;
                    const float h = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)row[xi[0]], wx[0]), __fmul_rn((float)row[xi[1]], wx[1])),
                                                        __fmul_rn((float)row[xi[2]], wx[2])), __fmul_rn((float)row[xi[3]], wx[3]));
                    acc = r == 0 ? __fmul_rn(h, wy[0]) : __fadd_rn(acc, __fmul_rn(h, wy[r]));
                }
                v = acc;
            }
        }
        out[i] = v;
    }
}

// F1 <- F1 conj(F2) / |F1 conj(F2)|  (mulSpectrums conjB, magSpectrums, divSpectrums with its FLT_EPSILON guard)
__global__ void cross_power_kernel(cufftComplex *__restrict__ f1, const cufftComplex *__restrict__ f2, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const cufftComplex p = f1[i], q = f2[i];
        const double re = (double)p.x * q.x + (double)p.y * q.y, im = (double)p.y * q.x - (double)p.x * q.y;
        const double mag = sqrt(re * re + im * im);
        const double s = mag / (mag * mag + (double)FLT_EPSILON);
        f1[i] = make_cuFloatComplex((float)(re * s), (float)(im * s));
    }
}

struct Peak {
    float v;
    unsigned long long idx; // index in the quadrant-swapped raster (minMaxLoc scans that one: first maximum wins)
};
__device__ __forceinline__ bool better(const Peak &a, const Peak &b) { return a.v > b.v || (a.v == b.v && a.idx < b.idx); }

__global__ void __launch_bounds__(256) peak_kernel(const float *__restrict__ c, int M, int N, Peak *__restrict__ out)
{
    __shared__ Peak s[8];
    Peak best{-FLT_MAX, ~0ull};
    const int64_t n = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / N), x = (int)(i - (int64_t)y * N);
        // OpenCV's fftShift = circular shift by (M/2, N/2) for even AND odd sizes (phasecorr.cpp: q0 lands at (xMid, yMid))
        const int ys = y + M / 2 >= M ? y + M / 2 - M : y + M / 2, xs = x + N / 2 >= N ? x + N / 2 - N : x + N / 2;
        const Peak p{c[i], (unsigned long long)ys * N + xs};
        if (better(p, best)) best = p;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        Peak p{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
        if (better(p, best)) best = p;
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (better(s[w], best)) best = s[w];
        out[blockIdx.x] = best;
    }
}

// final maximum + weightedCentroid (5x5, clamped to the image, double accumulators) -> {dx, dy, response}
__global__ void centroid_kernel(const float *__restrict__ c, int M, int N, const Peak *__restrict__ peaks, int n_peaks, double *out)
{
    if (threadIdx.x) return;
    Peak best = peaks[0];
    for (int i = 1; i < n_peaks; ++i)
        if (better(peaks[i], best)) best = peaks[i];
    const int py = (int)(best.idx / N), px = (int)(best.idx - (unsigned long long)py * N);
    const int y0 = max(py - 2, 0), y1 = min(py + 2, M - 1), x0 = max(px - 2, 0), x1 = min(px + 2, N - 1);
    double sx = 0.0, sy = 0.0, sum = 0.0;
    for (int ys = y0; ys <= y1; ++ys)
        for (int xs = x0; xs <= x1; ++xs) {
            const int y = ys >= M / 2 ? ys - M / 2 : ys + M - M / 2, x = xs >= N / 2 ? xs - N / 2 : xs + N - N / 2; // undo the shift
            const double v = (double)c[(int64_t)y * N + x];
            sx += xs * v;
            sy += ys * v;
            sum += v;
        }
    const double den = sum + DBL_EPSILON;
    out[0] = N / 2.0 - sx / den;
    out[1] = M / 2.0 - sy / den;
    out[2] = sum / ((double)M * N);
}

static const char *cufft_msg(cufftResult r)
{
    switch (r) {
    case CUFFT_SUCCESS: return "success";
    case CUFFT_ALLOC_FAILED: return "allocation failed";
    case CUFFT_INVALID_SIZE: return "invalid size";
    case CUFFT_INTERNAL_ERROR: return "internal error";
    case CUFFT_EXEC_FAILED: return "exec failed";
    case CUFFT_SETUP_FAILED: return "setup failed";
    default: return "error";
    }
}
#define OIP_CUFFT(call)                                                                              \
    do {                                                                                             \
        cufftResult r__ = (call);                                                                    \
        if (r__ != CUFFT_SUCCESS) return fail(OIP_E_CUDA, "%s: cuFFT %s (%d)", #call, cufft_msg(r__), (int)r__); \
    } while (0)

static int plans(oip_ctx *ctx, int M, int N, State **out)
{
    State *st = static_cast<State *>(ctx->stt_state);
    if (!st) ctx->stt_state = st = new State();
    if (!st->have || st->M != M || st->N != N) {
        if (st->have) { cufftDestroy(st->fwd); cufftDestroy(st->inv); st->have = false; }
        int n[2] = {M, N};
        OIP_CUFFT(cufftPlanMany(&st->fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, 2));
        OIP_CUFFT(cufftPlan2d(&st->inv, M, N, CUFFT_C2R));
        st->M = M; st->N = N; st->have = true;
    }
    OIP_CUFFT(cufftSetStream(st->fwd, ctx->stream));
    OIP_CUFFT(cufftSetStream(st->inv, ctx->stream));
    *out = st;
    return OIP_OK;
}

// one phase correlation, result left in d_out[3]; launches only (no host synchronisation).  pack(d_real, M, N, blocks)
// launches the kernel that fills the two zero-padded float planes.
template <typename Pack>
static int correlate_with(oip_ctx *ctx, int rows, int cols, double *d_out, Pack pack)
{
    const int M = optimal_dft_size(rows), N = optimal_dft_size(cols);
    State *st;
    int rc = plans(ctx, M, N, &st);
    if (rc) return rc;
    const int NC = N / 2 + 1;
    const size_t real_b = (size_t)2 * M * N * sizeof(float), cplx_b = (size_t)2 * M * NC * sizeof(cufftComplex);
    const int blocks = ctx->sm_count * 8;
    size_t o_c = (real_b + 255) & ~(size_t)255, o_p = (o_c + cplx_b + 255) & ~(size_t)255, total = o_p + (size_t)blocks * sizeof(Peak) + 256;
    rc = ensure_scratch(ctx, total);
    if (rc) return rc;
    uint8_t *S = (uint8_t *)ctx->d_scratch;
    float *d_real = (float *)S;
    cufftComplex *d_c = (cufftComplex *)(S + o_c);
    Peak *d_peaks = (Peak *)(S + o_p);
    pack(d_real, M, N, blocks);
    OIP_CUDA(cudaGetLastError());
    OIP_CUFFT(cufftExecR2C(st->fwd, d_real, d_c));
    cross_power_kernel<<<blocks, 256, 0, ctx->stream>>>(d_c, d_c + (size_t)M * NC, (int64_t)M * NC);
    OIP_CUDA(cudaGetLastError());
    OIP_CUFFT(cufftExecC2R(st->inv, d_c, d_real)); // unscaled, like cv::idft without DFT_SCALE
    peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d_real, M, N, d_peaks);
    OIP_CUDA(cudaGetLastError());
    centroid_kernel<<<1, 32, 0, ctx->stream>>>(d_real, M, N, d_peaks, blocks, d_out);
    OIP_CUDA(cudaGetLastError());
    ctx->launches += 4;
    return OIP_OK;
}

static int correlate(oip_ctx *ctx, const uint16_t *d_a, int64_t pitch_a, const uint16_t *d_b, int64_t pitch_b, int rows, int cols,
                     double *d_out)
{
    return correlate_with(ctx, rows, cols, d_out, [&](float *d_real, int M, int N, int blocks) {
        pack_kernel<<<blocks, 256, 0, ctx->stream>>>(d_a, pitch_a, d_b, pitch_b, rows, cols, M, N, d_real);
    });
}

// least-squares polynomial of degree deg (ascending coefficients), like Poly1d::fit (ref preproc.h:533-534): normal
// equations in long double on the centred / scaled abscissa, expanded back to powers of x
static bool polyfit(const std::vector<double> &x, const std::vector<double> &y, int deg, double *coef)
{
    const int n = (int)x.size(), m = deg + 1;
    if (n < m) return false;
    long double mean = 0, sc = 0;
    for (double v : x) mean += v;
    mean /= n;
    for (double v : x) sc = std::max(sc, fabsl((long double)v - mean));
    if (sc == 0) sc = 1;
    long double A[3][4] = {};
    for (int i = 0; i < n; ++i) {
        const long double t = ((long double)x[i] - mean) / sc;
        long double p[5] = {1, t, t * t, t * t * t, t * t * t * t};
        for (int r = 0; r < m; ++r) {
            for (int c = 0; c < m; ++c) A[r][c] += p[r + c];
            A[r][m] += p[r] * (long double)y[i];
        }
    }
    for (int c = 0; c < m; ++c) { // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < m; ++r)
            if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        if (fabsl(A[piv][c]) < 1e-300L) return false;
        for (int k = 0; k <= m; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = 0; r < m; ++r) {
            if (r == c) continue;
            const long double f = A[r][c] / A[c][c];
            for (int k = c; k <= m; ++k) A[r][k] -= f * A[c][k];
        }
    }
    long double q[3] = {0, 0, 0}; // coefficients in t
    for (int r = 0; r < m; ++r) q[r] = A[r][m] / A[r][r];
    // t = (x - mean) / sc  ->  powers of x
    const long double a = 1 / sc, b = -mean / sc;
    long double c0 = q[0] + q[1] * b + q[2] * b * b, c1 = q[1] * a + 2 * q[2] * a * b, c2 = q[2] * a * a;
    coef[0] = (double)c0;
    if (deg >= 1) coef[1] = (double)c1;
    if (deg >= 2) coef[2] = (double)c2;
    return true;
}

void destroy(oip_ctx *ctx)
{
    State *st = static_cast<State *>(ctx->stt_state);
    if (!st) return;
    if (st->have) { cufftDestroy(st->fwd); cufftDestroy(st->inv); }
    delete st;
    ctx->stt_state = nullptr;
}

} // namespace stt
} // namespace oip

using namespace oip;

extern "C" int oip_phase_correlate_u16(oip_ctx *ctx, const uint16_t *d_a, int64_t pitch_a_px, const uint16_t *d_b, int64_t pitch_b_px,
                                       int rows, int cols, double result[3])
{
    OIP_CHECK_CTX(ctx);
    if (!d_a || !d_b || !result) return fail(OIP_E_INVALID, "oip_phase_correlate_u16: null pointer");
    if (rows < 1 || cols < 1 || pitch_a_px < cols || pitch_b_px < cols) return fail(OIP_E_INVALID, "oip_phase_correlate_u16: bad geometry");
    int rc = ensure_pinned(ctx, 64);
    if (rc) return rc;
    double *d_out;
    OIP_CUDA(cudaMalloc(&d_out, 3 * sizeof(double)));
    rc = stt::correlate(ctx, d_a, pitch_a_px, d_b, pitch_b_px, rows, cols, d_out);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_pinned, d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(OIP_E_CUDA, "oip_phase_correlate_u16: %s", cudaGetErrorString(e));
        else memcpy(result, ctx->h_pinned, 3 * sizeof(double));
    }
    cudaFree(d_out);
    return rc;
}

extern "C" int oip_stt_parameters(oip_ctx *ctx, const uint16_t *d_pan1, const uint16_t *d_pan2, int w, int64_t total_lines,
                                  int64_t row0, int64_t rows_here, int64_t pitch_px, const oip_stt_config *cfg,
                                  oip_stt_section *sections_out, double sums[4])
{
    OIP_CHECK_CTX(ctx);
    if (sums) sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    if (!d_pan1 || !d_pan2 || !cfg || !sections_out) return fail(OIP_E_INVALID, "oip_stt_parameters: null pointer");
    const int ov = cfg->overlap_cols, ec = cfg->edge_cols, ns = cfg->sections, lps = cfg->lines_per_section;
    if (w < 1 || pitch_px < w || ov < 1 || ov > w || ec < 0 || ec >= ov || ns < 1 || lps < 1 || total_lines < (int64_t)ns * lps)
        return fail(OIP_E_INVALID, "oip_stt_parameters: bad geometry (w=%d overlap=%d edge=%d sections=%d x %d lines of %lld)", w, ov, ec,
                    ns, lps, (long long)total_lines);
    const int64_t gap = (total_lines - (int64_t)ns * lps) / (ns + 1); // ref stitcher.h:151
    const int64_t step = gap + lps;                                    // :152
    const int cols = ov - ec;                                          // :175-176 colRange(W-ov, W-ec) / colRange(ec, ov)
    int rc = ensure_pinned(ctx, 64 + (size_t)ns * 3 * sizeof(double));
    if (rc) return rc;
    double *d_out;
    OIP_CUDA(cudaMalloc(&d_out, (size_t)ns * 3 * sizeof(double)));
    OIP_CUDA(cudaMemsetAsync(d_out, 0, (size_t)ns * 3 * sizeof(double), ctx->stream));
    for (int i = 0; i < ns && !rc; ++i) {
        const int64_t off = gap + i * step;                            // :167
        sections_out[i].line_offset = off;
        sections_out[i].dx = sections_out[i].dy = sections_out[i].response = 0.0;
        sections_out[i].valid = -1;                                    // not held by this shard
        if (off < row0 || off + lps > row0 + rows_here) continue;
        const uint16_t *a = d_pan1 + (off - row0) * pitch_px + (w - ov);
        const uint16_t *b = d_pan2 + (off - row0) * pitch_px + ec;
        rc = stt::correlate(ctx, a, pitch_px, b, pitch_px, lps, cols, d_out + 3 * i);
        sections_out[i].valid = 0;
    }
    if (!rc) {
        double *h = (double *)((uint8_t *)ctx->h_pinned + 64);
        cudaError_t e = cudaMemcpyAsync(h, d_out, (size_t)ns * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(OIP_E_CUDA, "oip_stt_parameters: %s", cudaGetErrorString(e));
        for (int i = 0; i < ns && !rc; ++i) {
            oip_stt_section &s = sections_out[i];
            if (s.valid < 0) continue;
            s.dx = h[3 * i]; s.dy = h[3 * i + 1]; s.response = h[3 * i + 2];
            const bool ok = s.response >= cfg->threshold && (cfg->max_delta_y <= 0.0 || fabs(s.dy) <= cfg->max_delta_y); // :181
            s.valid = ok ? 1 : 0;
            if (ok && sums) { sums[0] += s.dx; sums[1] += s.dy; sums[2] += s.response; sums[3] += 1.0; } // :183-186
        }
    }
    cudaFree(d_out);
    return rc;
}

extern "C" int oip_inter_band_correlation(oip_ctx *ctx, const uint16_t *d_pan, int w, int64_t lines_pan, int64_t pan_pitch_px,
                                          const uint16_t *d_mss, int64_t lines_mss, int64_t mss_pitch_px, const oip_ibc_config *cfg,
                                          oip_ibc_shift *shifts, double cX[8], double cY[12])
{
    OIP_CHECK_CTX(ctx);
    if (!d_pan || !d_mss || !cfg || !shifts) return fail(OIP_E_INVALID, "oip_inter_band_correlation: null pointer");
    const int slices = cfg->slices, sections = cfg->sections, corr = cfg->correlation_lines > 0 ? cfg->correlation_lines : 16000;
    const int min_slices = cfg->min_slices > 0 ? cfg->min_slices : 8, min_count = cfg->min_count > 0 ? cfg->min_count : 5;
    if (slices < min_slices) return fail(OIP_E_INVALID, "CalcInterBandCorrelation: at lease %d slice needed", min_slices);           // ref preproc.h:228-230
    if (sections <= 0) return fail(OIP_E_INVALID, "CalcInterBandCorrelation: section count should be a positive integer");          // :231-233
    if (sections > 1 && (int64_t)sections * corr > lines_pan)                                                                       // :234-237
        return fail(OIP_E_INVALID, "CalcInterBandCorrelation: too many sections (%d lines per section), not enough total PAN data lines", corr);
    if (w < 4 * slices || (w & 3) || pan_pitch_px < w || mss_pitch_px < w || lines_pan < 4 || lines_mss < 1)
        return fail(OIP_E_INVALID, "oip_inter_band_correlation: bad geometry");
    const int base_rows = (int)std::min<int64_t>(lines_pan, corr);                     // :245
    const int gap = (int)((lines_pan - (int64_t)base_rows * sections) / (sections + 1)); // :246
    const int cols = w / slices;                                                        // :247
    const int brows = base_rows / 4, bgap = gap / 4, bcols = cols / 4, wb = w / 4;      // :272-274, band width = W / MSS_BANDS
    if (brows < 1 || bcols < 1) return fail(OIP_E_INVALID, "oip_inter_band_correlation: slices too small");
    const int total = slices * sections;
    int rc = ensure_pinned(ctx, 64 + (size_t)total * 4 * 3 * sizeof(double));
    if (rc) return rc;
    double *d_out;
    OIP_CUDA(cudaMalloc(&d_out, (size_t)total * 4 * 3 * sizeof(double)));
    for (int sec = 0; sec < sections && !rc; ++sec) {
        const int64_t r0 = gap + (int64_t)sec * (base_rows + gap);                      // :256
        const int64_t q0 = bgap + (int64_t)sec * (brows + bgap);                        // :283
        if (r0 < 0 || r0 + base_rows > lines_pan || q0 < 0 || q0 + brows > lines_mss) {
            rc = fail(OIP_E_RANGE, "oip_inter_band_correlation: section %d leaves the image (PAN %lld lines, MSS %lld lines)", sec,
                      (long long)lines_pan, (long long)lines_mss);
            break;
        }
        for (int i = 0; i < slices && !rc; ++i)
            for (int b = 0; b < 4 && !rc; ++b) {
                const uint16_t *pp = d_pan + r0 * pan_pitch_px + (int64_t)i * cols;
                const uint16_t *bp = d_mss + q0 * mss_pitch_px + (int64_t)b * wb + (int64_t)i * bcols; // band b = columns [b*wb, (b+1)*wb) (ref preproc.h:62-75)
                rc = stt::correlate_with(ctx, base_rows, cols, d_out + 3 * ((size_t)b * total + (size_t)sec * slices + i),
                                         [&](float *d_real, int M, int N, int blocks) {
                                             stt::pack_resize_kernel<<<blocks, 256, 0, ctx->stream>>>(
                                                 pp, pan_pitch_px, bp, mss_pitch_px, base_rows, cols, brows, bcols, (double)brows / base_rows,
                                                 (double)bcols / cols, M, N, d_real);
                                         });
            }
    }
    if (!rc) {
        double *h = (double *)((uint8_t *)ctx->h_pinned + 64);
        cudaError_t e = cudaMemcpyAsync(h, d_out, (size_t)total * 4 * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(OIP_E_CUDA, "oip_inter_band_correlation: %s", cudaGetErrorString(e));
        for (int b = 0; b < 4 && !rc; ++b) {
            std::vector<double> xs, dxs, dys;
            for (int k = 0; k < total; ++k) {
                oip_ibc_shift &s = shifts[(size_t)b * total + k];
                const double *v = h + 3 * ((size_t)b * total + k);
                s.dx = v[0]; s.dy = v[1]; s.rs = v[2];
                s.cx = (k % slices) * cols + cols / 2;                                  // :326
                s.pad = 0;
                if (s.rs < cfg->threshold) { s.dx = s.dy = std::nan(""); continue; }     // FilterInterBandShiftValues :495-503
                xs.push_back((double)s.cx); dxs.push_back(s.dx); dys.push_back(s.dy);
            }
            if ((int)xs.size() < min_count) {                                           // :505-510
                rc = fail(OIP_E_RANGE, "Not enough valid correlation values for band#%d: %d valid values found, %d expected at least", b + 1,
                          (int)xs.size(), min_count);
                break;
            }
            if (cX && cY) {                                                             // DoCorrelationPolynomialFitting :513-547
                if (!stt::polyfit(xs, dxs, 1, cX + 2 * b) || !stt::polyfit(xs, dys, 2, cY + 3 * b))
                    rc = fail(OIP_E_RANGE, "polynomial fit of band#%d is singular", b + 1);
            }
        }
    }
    cudaFree(d_out);
    return rc;
}
