/* TEST INFRASTRUCTURE (oracle/_ref build only). Force-included ahead of the reference's
 * NN translation units: keeps ARM_OPTIMIZED 1 so affine*.c read the shipped weight tables
 * in the 4-row interleave that python/nnsp_pack/c_weight_man.py writes. */
#define __AMBIQ_NNSP_DEBUG__
#define AMBIQ_NNSP_DEBUG 0
#define ARM_OPTIMIZED 1
