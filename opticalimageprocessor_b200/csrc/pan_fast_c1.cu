// pan_fast_c1.cu -- pan_fast_kernel for sub-image tiles of the IMDT stream (OIP_FMT_BE16_TILES)
#define OIP_FAST_CLS 1
#include "pan_fast_dev.cuh"
