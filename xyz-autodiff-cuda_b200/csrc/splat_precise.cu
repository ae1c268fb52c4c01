// csrc/splat_precise.cu -- IEEE flavour of the splat kernels (built without fast-math flags).
#define XYZ_SPLAT_FLAVOR precise
#define XYZ_SPLAT_IS_FAST 0
#include "splat_kernels.cuh"
