/*
 * nnsp_legacy_api.h -- the ns-nnsp C interface that nnsp-b200 stays binary compatible with.
 *
 * This is an independent declaration of the *interface* (type names, field order, function
 * signatures) of the reference library, so that existing callers and the generated model
 * tables (def_nn<id>_<name>.c) compile and link against libnnsp_b200.so unchanged.
 * The implementation behind every symbol is the CUDA engine in nnsp_b200/csrc.
 *
 * Interface provenance (reference file:line, AmbiqAI/nnsp):
 *   ACTIVATION_TYPE ............ ns-nnsp/includes-api/activation.h:16-22
 *   NET_LAYER_TYPE ............. ns-nnsp/includes-api/neural_nets.h:9-13
 *   NeuralNetClass ............. ns-nnsp/includes-api/neural_nets.h:15-32
 *   stftModule ................. ns-nnsp/includes-api/spectrogram_module.h:9-16
 *   FeatureClass ............... ns-nnsp/includes-api/feature_module.h:7-18
 *   NNSPClass .................. ns-nnsp/includes-api/nn_speech.h:12-25
 *   NNSP_ID .................... ns-nnsp/includes-api/nnsp_identification.h:3-9
 *   constants .................. ns-nnsp/includes-api/ambiq_nnsp_const.h:3-10, s2i_const.h:3-4
 * tests/test_abi.py checks sizeof/offsetof of every struct against the reference
 * headers whenever /root/reference is present.
 */
#ifndef NNSP_B200_LEGACY_API_H
#define NNSP_B200_LEGACY_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants of the front end --------------------------------------------------- */
#define LEN_FFT_NNSP          512
#define LEN_STFT_WIN_COEFF    480
#define LEN_STFT_HOP          160
#define NUM_MELBANKS          40
#define NUM_FEATURE_CONTEXT   6
#define MAX_SIZE_FEATURE      50
#define DIMEMSION_FEATURE     NUM_MELBANKS
#define SAMPLING_RATE         16000
#define DIM_INTENTS           7
#define DIM_SLOTS             17

/* helper macros some callers pick up from the library's public headers (minmax.h, ambiq_stdint.h) */
#ifndef MAX
#define MAX(x, y) (((x) > (y)) ? (x) : (y))
#endif
#ifndef MIN
#define MIN(x, y) (((x) < (y)) ? (x) : (y))
#endif
#ifndef MAX_INT32_T
#define MAX_INT32_T ((int32_t)0x7fffffff)
#define MIN_INT32_T ((int32_t)0x80000000)
#define MAX_INT16_T ((int16_t)0x7fff)
#define MIN_INT16_T ((int16_t)0x8000)
#define ONE_N32_Q15 ((int32_t)32768)
#endif

typedef enum { s2i_id = 0, vad_id = 1, kws_galaxy_id = 2, num_NNSP_IDS = 3 } NNSP_ID;

/* ---- neural net description (what def_nn*.c instantiates) ------------------------- */
typedef enum { relu6, ftanh, sigmoid, linear } ACTIVATION_TYPE;
typedef enum { fc, lstm } NET_LAYER_TYPE;

typedef struct {
    int8_t          numlayers;
    int16_t         size_layer[11];        /* [0] = input width, [i+1] = width of layer i */
    NET_LAYER_TYPE  net_layer_type[10];
    int8_t          qbit_kernel[10];
    int8_t          qbit_input[10];
    int8_t          qbit_bias[10];
    ACTIVATION_TYPE activation_type[10];
    int32_t        *pt_cstate[10];
    int16_t        *pt_hstate[10];
    void *(*act_func[10])(void *, int32_t *, int);
    int  *(*layer_func[10])();
    int8_t         *pt_kernel[10];
    int16_t        *pt_bias[10];
    int8_t         *pt_kernel_rec[10];
} NeuralNetClass;

void NeuralNetClass_init(NeuralNetClass *pt_inst);
void NeuralNetClass_setDefault(NeuralNetClass *pt_inst);
void NeuralNetClass_exe(NeuralNetClass *pt_inst, int16_t *input, int32_t *output,
                        int8_t debug_layer);

/* Activation / layer entry points. The model tables store their addresses in
 * act_func[] / layer_func[]; nnsp-b200 reads those addresses to learn each layer's
 * activation and accumulator width (64-bit vs ACC32BIT_OPT wrapping 32-bit). */
void *relu6_fix(int16_t *y, int32_t *x, int len);
void *linear_fix(int32_t *y, int32_t *x, int len);
void *tanh_fix(int16_t *y, int32_t *x, int len);
void *sigmoid_fix(int16_t *y, int32_t *x, int len);

#define NNSP_LEGACY_LAYER_ARGS                                                         \
    int16_t *p_output, int8_t *p_kernel, int8_t *p_kernel_rec, int16_t *p_bias,        \
    int16_t *input, int16_t *input_rec, int32_t *c_state, int16_t dim_output,          \
    int16_t dim_input, int16_t dim_input_rec, int16_t qbit_kernel, int16_t qbit_bias,  \
    int16_t qbit_input, int16_t qbit_input_rec, ACTIVATION_TYPE act_type,              \
    void *(*act)(void *, int32_t *, int)
int fc_8x16(NNSP_LEGACY_LAYER_ARGS);
int fc_8x16_acc32b(NNSP_LEGACY_LAYER_ARGS);
int lstm_8x16(NNSP_LEGACY_LAYER_ARGS);
int lstm_8x16_acc32b(NNSP_LEGACY_LAYER_ARGS);

/* ---- feature front end ------------------------------------------------------------ */
typedef struct {
    int16_t        len_win;
    int16_t        hop;
    int16_t        len_fft;
    int16_t        dataBuffer[512];
    const int16_t *window;
} stftModule;

typedef struct {
    stftModule     state_stftModule;
    int32_t        feature[MAX_SIZE_FEATURE];
    int16_t        normFeatContext[NUM_FEATURE_CONTEXT * MAX_SIZE_FEATURE];
    int16_t        num_context;
    int16_t        dim_feat;
    const int32_t *pt_norm_mean;
    const int32_t *pt_norm_stdR;
    int8_t         qbit_output;
} FeatureClass;

void FeatureClass_construct(FeatureClass *ps, const int32_t *norm_mean,
                            const int32_t *norm_stdR, int8_t qbit_output);
void FeatureClass_setDefault(FeatureClass *ps);
void FeatureClass_execute(FeatureClass *ps, int16_t *input);

/* ---- per-frame driver --------------------------------------------------------------- */
typedef struct {
    char     nn_id;
    void    *pt_net;                 /* NeuralNetClass* */
    void    *pt_feat;                /* FeatureClass*   */
    int8_t   slides;
    int16_t  trigger;
    int16_t *pt_thresh_prob;
    int16_t  counts_category[8];
    int16_t *pt_th_count_trigger;
    int16_t  num_dnsmpl;
    int16_t  outputs[3];
    int16_t  argmax_last;
} NNSPClass;

int NNSPClass_init(NNSPClass *pt_inst, void *pt_net, void *pt_feat, char nn_id,
                   const int32_t *pt_mean, const int32_t *pt_stdR,
                   int16_t *pt_thresh_prob, int16_t *pt_th_count_trigger);
int NNSPClass_reset(NNSPClass *pt_inst);
int16_t NNSPClass_exec(NNSPClass *pt_inst, int16_t *rawPCM);

#ifdef __cplusplus
}
#endif
#endif /* NNSP_B200_LEGACY_API_H */
