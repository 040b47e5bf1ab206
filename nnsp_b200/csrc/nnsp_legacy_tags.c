/* nnsp_legacy_tags.c -- the layer/activation entry points whose ADDRESSES the reference model
 * tables store in NeuralNetClass.layer_func[] / act_func[] (evb/src/def_nn1_vad.c:61-84).
 *
 * In nnsp-b200 the per-layer arithmetic runs inside the fused CUDA network kernel
 * (nnsp_engine.cu, nnsp_split.cu, nnsp_compat.cu), never one layer at a time on the host, so these symbols exist to be
 * compared against (nnsp_model_net.c) and to let unmodified def_nn*.c objects link. Calling
 * one directly is a usage error: they report it and compute nothing (no CPU fallback). */
#include "nnsp_compat/nnsp_legacy_api.h"
#include "nnsp_model.h"

#define TAG_BODY(name)                                                                      \
    nnsp_set_error(name " is an identity tag in nnsp-b200; run the network through "        \
                        "NeuralNetClass_exe / nnsp_b200_batch_exec");                       \
    return 0

void *relu6_fix(int16_t *y, int32_t *x, int len)   { (void)y; (void)x; (void)len; TAG_BODY("relu6_fix"); }
void *linear_fix(int32_t *y, int32_t *x, int len)  { (void)y; (void)x; (void)len; TAG_BODY("linear_fix"); }
void *tanh_fix(int16_t *y, int32_t *x, int len)    { (void)y; (void)x; (void)len; TAG_BODY("tanh_fix"); }
void *sigmoid_fix(int16_t *y, int32_t *x, int len) { (void)y; (void)x; (void)len; TAG_BODY("sigmoid_fix"); }

#define LAYER_TAG(name)                                                                     \
    int name(NNSP_LEGACY_LAYER_ARGS)                                                        \
    {                                                                                       \
        (void)p_output; (void)p_kernel; (void)p_kernel_rec; (void)p_bias; (void)input;      \
        (void)input_rec; (void)c_state; (void)dim_output; (void)dim_input;                  \
        (void)dim_input_rec; (void)qbit_kernel; (void)qbit_bias; (void)qbit_input;          \
        (void)qbit_input_rec; (void)act_type; (void)act;                                    \
        nnsp_set_error(#name " is an identity tag in nnsp-b200; run the network through "   \
                             "NeuralNetClass_exe / nnsp_b200_batch_exec");                  \
        return -1;                                                                          \
    }
LAYER_TAG(fc_8x16)
LAYER_TAG(fc_8x16_acc32b)
LAYER_TAG(lstm_8x16)
LAYER_TAG(lstm_8x16_acc32b)
