// pan_fast.cu -- the regular-interior fast path of the fused PAN pipeline (unpack -> RRC -> cubic shift -> concat).
//
// Everything the planner can prove regular (footprints inside the section's fresh rows and the
// CCD's columns, fixed-point map advancing by exactly one source pixel per output pixel) runs here; section
// edges, image borders, map-rounding anomalies, packed / tiled / unaligned inputs stay on the exact generic
// kernel of pan_pipeline.cu.  Replaces the same reference code as oip_pan_pipeline (ref imageop.h:129-138,
// :230-275, :277-363, stitcher.h:83-139).
//
// Structure: a CTA is four INDEPENDENT warps, each marching down its own warp-tile; there is no shared ring
// and no __syncthreads.
//   load      one lane issues 2-D TMA tensor copies (cp.async.bulk.tensor, SASS UTMALDG) of RC raw rows into
//             the warp's private stage ring, n_stage deep, each stage behind its own mbarrier
//   convert   a lane owns 4 + 4 source columns (left half / right half of the tile): byte swap by PRMT, fp64
//             RRC with (k,b) in registers (magic-number int<->double, no conversion-pipe op), one I2F.U16
//   exchange  the 3 extra window columns of each half come from lane+1 by SHFL (no shared-memory ring)
//   resample  scatter form of OpenCV's bicubic sum: a source row is multiplied once into the 4 output rows
//             it feeds (accumulators rotate through registers), products and sums in OpenCV's own order,
//             packed FFMA2/FADD2 over the (left, right) pixel pair, no FMA contraction
//   store     F2I.U16 (round-half-even, saturating) + 64-bit stores straight into the trimmed output raster
// COPY warp-tiles (unshifted CCDs) use the same staging: swap + RRC + 128-bit stores.
// The device code lives in pan_fast_dev.cuh and is compiled once per source-format class (pan_fast_c0/_c1/_c2.cu); this
// file holds the host side: tensor-map encoding and the launches.
#include "pan_fast.cuh"
#include "tma_warp.cuh"

namespace oip {

namespace tmaw {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int encode_tmap_u32(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes, int box_w32, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)(w / 2), (cuuint64_t)n_rows}; // 32-bit elements = sample pairs
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OIP_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return OIP_OK;
}
} // namespace tmaw

namespace panfast {
int fast_encode_tmap(CUtensorMap *tm, const void *base, int w, int64_t n_rows, int64_t pitch_bytes)
{
    return tmaw::encode_tmap_u32(tm, base, w, n_rows, pitch_bytes, BOX_W, RC);
}

int fast_encode_tmap_packed(CUtensorMap *tm, const void *base, int w, int bits, int64_t n_rows, int64_t pitch_bytes)
{
    // the packed line as 32-bit elements (w * bits / 32 of them): encode_tmap_u32 counts w / 2 elements for w samples
    return tmaw::encode_tmap_u32(tm, base, 2 * (w * bits / 32), n_rows, pitch_bytes, bits == 12 ? PBOX12 : PBOX10, RC);
}

// cls_ctas[c]: CTAs of class c; the FastTile array holds the classes one after the other (each padded to whole CTAs)
int fast_launch(oip_ctx *ctx, const FastParams &P, const int64_t *cls_ctas)
{
    int64_t tile0 = 0;
    for (int c = 0; c < 3; ++c) {
        if (cls_ctas[c] <= 0) continue;
        int rc = c == 0 ? fast_launch_cls<0>(ctx, P, tile0, cls_ctas[c]) : c == 1 ? fast_launch_cls<1>(ctx, P, tile0, cls_ctas[c]) : fast_launch_cls<2>(ctx, P, tile0, cls_ctas[c]);
        if (rc) return rc;
        tile0 += cls_ctas[c] * WARPS;
    }
    return OIP_OK;
}

} // namespace panfast
} // namespace oip
