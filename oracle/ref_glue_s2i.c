/* oracle/ref_glue_s2i.c -- TEST INFRASTRUCTURE, never linked into or called by the product.
 *
 * Driver around the reference's S2I-only controller, evb/src/s2iCntrlClass.c, compiled UNMODIFIED into
 * oracle/_ref/libnnsp_ref_s2ictrl.so (its own library: it defines `intents`, `slots` and a ParamCntrlClass of its
 * own, which clash with nnCntrlClass.c). It pins the claim that `nnsp_b200_cascade_create` with seq = {s2i_id} and
 * frs_vbufBk_s2i = 0 is that controller: tests/test_oracle_vs_ref.py (CPU restatement) and
 * tests/test_gpu_cascade.py (CUDA engine) compare frame by frame with what this returns. */
#include <stdarg.h>
#include <stdint.h>
#include <string.h>

#include "nn_speech.h"
#include "feature_module.h"
#include "neural_nets.h"
#include "PcmBufClass.h"
#include "s2iCntrlClass.h"

typedef struct { int8_t stage_id, pos_after; int16_t detected; int16_t outputs[3];
                 uint16_t cnt_timeout; } glue_cascade_result;                   /* == nnsp_b200_cascade_result */

extern NNSPClass NNSP_INST;
extern FeatureClass FEAT_INST;
extern const char *intents[];
extern const char *slots[];

static int g_evt_detect;
static int16_t g_evt_outputs[3];

/* oracle/ref_shim/ns_ambiqsuite_harness.h routes ns_lp_printf here: the detection message (s2iCntrlClass.c:116-119)
 * is the only place the outputs are visible before the controller resets the instance */
void ref_glue_printf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (strstr(fmt, "Detected: %s")) {
        const char *a = va_arg(ap, const char *), *b = va_arg(ap, const char *), *c = va_arg(ap, const char *);
        int i;
        g_evt_detect = 1;
        for (i = 0; i < 7; i++) if (intents[i] == a) g_evt_outputs[0] = (int16_t)i;
        for (i = 0; i < 17; i++) { if (slots[i] == b) g_evt_outputs[1] = (int16_t)i; if (slots[i] == c) g_evt_outputs[2] = (int16_t)i; }
    }
    va_end(ap);
}

static s2iCntrlClass g_ctl;

/* do_reset: 1 = brand-new controller (zeroed instances, init, reset), 2 = s2iCntrlClass_reset of the live one, 0 = go on.
 * thresh_prob / thresh_cnts < 0 keep the defaults of ParamsNNCntrl.h. cnt_timeout is not a notion of this controller: 0. */
int ref_s2ictrl_run(int do_reset, int thresh_prob, int thresh_cnts, const int16_t *pcm, int n_frames,
                    glue_cascade_result *results)
{
    static int16_t frame[160];
    int t, i;
    if (do_reset == 1) {
        memset(&NNSP_INST, 0, sizeof NNSP_INST);
        memset(&FEAT_INST, 0, sizeof FEAT_INST);
        s2iCntrlClass_init(&g_ctl);
        if (thresh_prob >= 0) g_ctl.Params.thresh_prob_s2i = (int16_t)thresh_prob;
        if (thresh_cnts >= 0) g_ctl.Params.thresh_cnts_s2i = (int16_t)thresh_cnts;
        s2iCntrlClass_reset(&g_ctl);
    } else if (do_reset == 2) {
        s2iCntrlClass_reset(&g_ctl);
    }
    for (t = 0; t < n_frames; t++) {
        memcpy(frame, pcm + (size_t)t * 160, sizeof frame);
        g_evt_detect = 0; g_evt_outputs[0] = g_evt_outputs[1] = g_evt_outputs[2] = 0;
        s2iCntrlClass_exec(&g_ctl, frame);
        if (results) {
            results[t].stage_id = 0; results[t].pos_after = 0;
            results[t].detected = (int16_t)g_evt_detect;
            for (i = 0; i < 3; i++) results[t].outputs[i] = g_evt_outputs[i];
            results[t].cnt_timeout = 0;
        }
    }
    return 0;
}
