/* nnsp_model_net.c -- read a live reference model table (`NeuralNetClass`, as instantiated by
 * evb/src/def_nn0_s2i.c:29-110 and friends) into the canonical nnsp_b200_model.
 * Product code only (the oracle loads blobs). */
#include "nnsp_model.h"
#include "nnsp_compat/nnsp_legacy_api.h"
#include <stdlib.h>
#include <string.h>

int nnsp_model_alloc_layer(nnsp_layer *L);
void nnsp_model_layer_from_table(nnsp_layer *L, const int8_t *kernel, const int8_t *kernel_rec,
                                 const int16_t *bias);
int nnsp_model_validate(const struct nnsp_b200_model *m);

static int act_from_table(const NeuralNetClass *n, int i)
{
    /* the reference computes with act_func[i] (neural_nets.c:114) and consults
     * activation_type[i] only for the final int32 copy (:152); prefer the pointer when it is
     * one of this library's own tags, otherwise trust the enum */
    void *(*f)(void *, int32_t *, int) = n->act_func[i];
    if (f == (void *(*)(void *, int32_t *, int))&tanh_fix) return NNSP_ACT_TANH;
    if (f == (void *(*)(void *, int32_t *, int))&sigmoid_fix) return NNSP_ACT_SIGMOID;
    if (f == (void *(*)(void *, int32_t *, int))&relu6_fix) return NNSP_ACT_RELU6;
    if (f == (void *(*)(void *, int32_t *, int))&linear_fix) return NNSP_ACT_LINEAR;
    switch (n->activation_type[i]) {
    case relu6: return NNSP_ACT_RELU6;
    case ftanh: return NNSP_ACT_TANH;
    case sigmoid: return NNSP_ACT_SIGMOID;
    case linear: return NNSP_ACT_LINEAR;
    }
    return -1;
}

int nnsp_b200_model_from_net(const void *neural_net_class, const int32_t *feature_mean,
                             const int32_t *feature_stdR, int nn_id, nnsp_b200_model **out)
{
    const NeuralNetClass *n = (const NeuralNetClass *)neural_net_class;
    if (!n || !feature_mean || !feature_stdR || !out) return NNSP_B200_ERR_ARG;
    if (n->numlayers < 1 || n->numlayers > NNSP_B200_MAX_LAYERS) {
        nnsp_set_error("numlayers %d outside 1..%d", (int)n->numlayers, NNSP_B200_MAX_LAYERS);
        return NNSP_B200_ERR_ARG;
    }
    nnsp_b200_model *m = (nnsp_b200_model *)calloc(1, sizeof *m);
    if (!m) return NNSP_B200_ERR_NOMEM;
    m->nn_id = nn_id;
    m->numlayers = n->numlayers;
    memcpy(m->size_layer, n->size_layer, sizeof m->size_layer);
    memcpy(m->mean, feature_mean, sizeof m->mean);
    memcpy(m->stdR, feature_stdR, sizeof m->stdR);
    int rc = NNSP_B200_OK;
    for (int i = 0; i < n->numlayers; i++) {
        nnsp_layer *L = &m->layer[i];
        L->type = (n->net_layer_type[i] == lstm) ? NNSP_LAYER_LSTM : NNSP_LAYER_FC;
        L->act = act_from_table(n, i);
        L->rows = n->size_layer[i + 1];
        L->cols = n->size_layer[i];
        L->qk = n->qbit_kernel[i];
        L->qi = n->qbit_input[i];
        L->qb = n->qbit_bias[i];
        /* neural_nets.c:108 reads qbit_input[i+1]; for i == 9 that is the byte after the array,
         * which the struct layout makes qbit_bias[0] */
        L->qi_next = (i + 1 < NNSP_B200_MAX_LAYERS) ? n->qbit_input[i + 1] : n->qbit_bias[0];
        int *(*lf)() = n->layer_func[i];
        L->acc32 = (lf == (int *(*)())&fc_8x16_acc32b || lf == (int *(*)())&lstm_8x16_acc32b);
        if (L->rows < 1 || L->cols < 1 || !n->pt_kernel[i] || !n->pt_bias[i] ||
            (L->type == NNSP_LAYER_LSTM && !n->pt_kernel_rec[i])) {
            nnsp_set_error("layer %d: missing kernel/bias table or empty shape", i);
            rc = NNSP_B200_ERR_ARG;
            break;
        }
        rc = nnsp_model_alloc_layer(L);
        if (rc) break;
        nnsp_model_layer_from_table(L, n->pt_kernel[i], n->pt_kernel_rec[i], n->pt_bias[i]);
    }
    if (rc == NNSP_B200_OK) rc = nnsp_model_validate(m);
    if (rc != NNSP_B200_OK) { nnsp_b200_model_free(m); return rc; }
    *out = m;
    return NNSP_B200_OK;
}
