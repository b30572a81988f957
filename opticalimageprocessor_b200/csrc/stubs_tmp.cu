// temporary: entry points not implemented yet
#include "oip_common.cuh"
using namespace oip;
extern "C" {
int oip_crc16_batch(oip_ctx *, const uint8_t *, const uint64_t *, int64_t, int, uint16_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_aos_scan(oip_ctx *, const uint8_t *, size_t, uint64_t *, size_t, int64_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_imtr_deframe(oip_ctx *, const uint8_t *, const uint64_t *, int64_t, uint8_t *, size_t, int64_t *, int64_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_image_frames_index(oip_ctx *, const uint8_t *, size_t, const oip_frame_geom *, oip_frame_entry *, int64_t, int64_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_unpack_frames(oip_ctx *, const uint8_t *, size_t, const oip_frame_geom *, const oip_frame_entry *, int64_t, uint8_t *, uint16_t *, uint16_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_band_align_merge(oip_ctx *, const void *, const oip_mss_desc *, uint16_t *, int64_t *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
int oip_pan_pipeline_host(oip_ctx *, const oip_pan_desc *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
}
