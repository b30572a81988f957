// temporary: entry points not implemented yet
#include "oip_common.cuh"
using namespace oip;
extern "C" {
int oip_pan_pipeline_host(oip_ctx *, const oip_pan_desc *) { return fail(OIP_E_UNSUPPORTED, "not implemented yet"); }
}
