/* TEST INFRASTRUCTURE (oracle/_ref build only). Force-included (-include) ahead of the
 * reference's feature/FFT translation units: pre-defines the include guard of
 * ns-nnsp/includes-api/ambiq_nnsp_debug.h so that its un-guarded `#define ARM_OPTIMIZED 1`
 * (line 4) is skipped and the in-tree fixed-point rfft (fft.c) is compiled instead of the
 * CMSIS-DSP wrapper whose source is not in the tree. */
#define __AMBIQ_NNSP_DEBUG__
#define AMBIQ_NNSP_DEBUG 0
#define ARM_OPTIMIZED 0
