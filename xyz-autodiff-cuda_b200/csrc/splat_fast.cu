// csrc/splat_fast.cu -- fast-math flavour of the splat kernels (built with -use_fast_math).
#define XYZ_SPLAT_FLAVOR fast
#define XYZ_SPLAT_IS_FAST 1
#include "splat_kernels.cuh"
