// pan_fast_c0.cu -- pan_fast_kernel for line rasters (LE16 / BE16)
#define OIP_FAST_CLS 0
#include "pan_fast_dev.cuh"
