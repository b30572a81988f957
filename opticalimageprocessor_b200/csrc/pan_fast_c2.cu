// pan_fast_c2.cu -- pan_fast_kernel for MSB-first packed lines (12 / 10 bit)
#define OIP_FAST_CLS 2
#include "pan_fast_dev.cuh"
