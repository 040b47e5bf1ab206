/* oracle/ref_glue.c -- TEST INFRASTRUCTURE, never linked into or called by the product.
 *
 * Thin driver around the UNMODIFIED reference sources compiled by oracle/Makefile into
 * oracle/_ref/libnnsp_ref_acc{64,32}.so. It only (a) feeds frames to the reference's own
 * entry points and (b) copies the reference's own state out after every frame, in the tap
 * layout of include/nnsp_b200.h, so that the CUDA engine and the C restatement
 * (oracle/nnsp_oracle.c) can be compared with it record by record.
 *
 * The reference is single-instance (global scratch): one stream at a
 * time, and one process per concurrent user.
 */
#include <stdarg.h>
#include <stdint.h>
#include <string.h>

#include "nn_speech.h"
#include "feature_module.h"
#include "neural_nets.h"
#include "nnsp_identification.h"
#include "affine.h"
#include "affine_acc32b.h"
#include "lstm.h"
#include "activation.h"
#ifndef GLUE_DROPIN          /* internals of the reference's front end: not part of the public interface */
#include "spectrogram_module.h"
#include "melSpecProc.h"
#include "fixlog10.h"
#include "fft.h"
#endif
#include "def_nn0_s2i.h"
#include "def_nn1_vad.h"
#include "def_nn2_kws_galaxy.h"
#include "PcmBufClass.h"
#include "nnCntrlClass.h"

typedef struct { int16_t trigger; int16_t outputs[3]; } glue_result;            /* == nnsp_b200_result */
typedef struct { int8_t stage_id, pos_after; int16_t detected; int16_t outputs[3];
                 uint16_t cnt_timeout; } glue_cascade_result;                   /* == nnsp_b200_cascade_result */

#ifndef GLUE_DROPIN
extern const int16_t stft_win_coeff[];
extern const int32_t fft_tw_coeff[], rfft_tw_coeff[];
extern const int16_t br_coeff[], mfltrBank_coeff[], log_tayler_coeff[];
extern int16_t coeffs_tanh[];

#endif

int ref_is_acc32(void)
{
#ifdef DEF_ACC32BIT_OPT
    return 1;
#else
    return 0;
#endif
}

int ref_model(int nn_id, void **net, const int32_t **mean, const int32_t **stdR)
{
    switch (nn_id) {
    case s2i_id:        *net = &net_s2i;        *mean = feature_mean_s2i;        *stdR = feature_stdR_s2i;        return 0;
    case vad_id:        *net = &net_vad;        *mean = feature_mean_vad;        *stdR = feature_stdR_vad;        return 0;
    case kws_galaxy_id: *net = &net_kws_galaxy; *mean = feature_mean_kws_galaxy; *stdR = feature_stdR_kws_galaxy; return 0;
    }
    return -1;
}

#ifndef GLUE_DROPIN
/* the reference's constant tables, for tests/test_tables_and_format.py */
int ref_table(const char *name, const void **p, int *elem_bytes)
{
    if (!strcmp(name, "stft_win")) { *p = stft_win_coeff;   *elem_bytes = 2; return 480; }
    if (!strcmp(name, "fft_tw"))   { *p = fft_tw_coeff;     *elem_bytes = 4; return 256; }
    if (!strcmp(name, "rfft_tw"))  { *p = rfft_tw_coeff;    *elem_bytes = 4; return 256; }
    if (!strcmp(name, "bitrev"))   { *p = br_coeff;         *elem_bytes = 2; return 256; }
    if (!strcmp(name, "mel"))      { *p = mfltrBank_coeff;  *elem_bytes = 2; return 534; }
    if (!strcmp(name, "log_lut"))  { *p = log_tayler_coeff; *elem_bytes = 2; return 256; }
    if (!strcmp(name, "tanh_lut")) { *p = coeffs_tanh;      *elem_bytes = 2; return 384; }
    return -1;
}

/* ---- front end, stage by stage, on one 480-sample analysis window --------------------- */
void ref_feature_stages(const int16_t *win480, int32_t *fft_in, int32_t *spec, int32_t *pspec,
                        int32_t *mel, int32_t *logmel)
{
    static stftModule st;
    static int32_t sp[1028], ps[1028], me[MAX_SIZE_FEATURE];
    int i;
    stftModule_construct(&st);
    /* stftModule_analyze slides dataBuffer by one hop and appends x: preload so the analysis
     * buffer equals win480 (spectrogram_module.c:55-60) */
    for (i = 0; i < 320; i++) st.dataBuffer[160 + i] = win480[i];
    stftModule_analyze(&st, (int16_t *)(win480 + 320), sp);
    if (fft_in) {   /* fft_in is function-static in the reference; same expression, reference window table */
        for (i = 0; i < 480; i++) fft_in[i] = ((int32_t)stft_win_coeff[i] * (int32_t)win480[i]) >> 15;
        for (i = 480; i < 512; i++) fft_in[i] = 0;
    }
    if (spec) memcpy(spec, sp, 514 * sizeof(int32_t));
    spec2pspec(ps, sp, 257);
    if (pspec) memcpy(pspec, ps, 257 * sizeof(int32_t));
    melSpecProc(ps, me);
    if (mel) memcpy(mel, me, 40 * sizeof(int32_t));
    log10_vec(me, me, 40, 15);
    if (logmel) memcpy(logmel, me, 40 * sizeof(int32_t));
}

#endif /* !GLUE_DROPIN */

/* ---- shared tap helpers ----------------------------------------------------------------- */
static int act_stride(const NeuralNetClass *n)
{
    int i, s = 0;
    for (i = 1; i < n->numlayers; i++) s += n->size_layer[i];
    return s;
}
static int h_stride(const NeuralNetClass *n)
{
    int i, s = 0;
    for (i = 0; i < n->numlayers; i++) if (n->net_layer_type[i] == lstm) s += n->size_layer[i + 1];
    return s;
}
int ref_strides(int nn_id, int *act, int *h, int *n_out)
{
    void *net; const int32_t *m, *s;
    if (ref_model(nn_id, &net, &m, &s)) return -1;
    NeuralNetClass *n = (NeuralNetClass *)net;
    *act = act_stride(n); *h = h_stride(n); *n_out = n->size_layer[n->numlayers];
    return 0;
}

static void save_hc(NeuralNetClass *n, int16_t *h, int32_t *c)
{
    int i, j, o = 0;
    for (i = 0; i < n->numlayers; i++)
        if (n->net_layer_type[i] == lstm)
            for (j = 0; j < n->size_layer[i + 1]; j++, o++) { h[o] = n->pt_hstate[i][j]; c[o] = n->pt_cstate[i][j]; }
}
static void load_hc(NeuralNetClass *n, const int16_t *h, const int32_t *c)
{
    int i, j, o = 0;
    for (i = 0; i < n->numlayers; i++)
        if (n->net_layer_type[i] == lstm)
            for (j = 0; j < n->size_layer[i + 1]; j++, o++) { n->pt_hstate[i][j] = h[o]; n->pt_cstate[i][j] = c[o]; }
}

/* Shadow evaluation: with the LSTM state temporarily set to (h0,c0), run the first k layers
 * for k = 1..L via the reference's own debug_layer tap (neural_nets.c:65,152-167) and record
 * every layer output; the live state is put back afterwards. */
static void shadow_layers(NeuralNetClass *n, int16_t *ctx, const int16_t *h0, const int32_t *c0,
                          int16_t *act, int32_t *logits)
{
    static int16_t hl[1024]; static int32_t cl[1024];
    static int32_t out[512];
    int k, o = 0, j;
    save_hc(n, hl, cl);
    for (k = 1; k <= n->numlayers; k++) {
        load_hc(n, h0, c0);
        NeuralNetClass_exe(n, ctx, out, (int8_t)k);
        if (k < n->numlayers) {
            if (act) memcpy(act + o, out, n->size_layer[k] * sizeof(int16_t));
            o += n->size_layer[k];
        } else if (logits) {
            if (n->activation_type[k - 1] == linear) memcpy(logits, out, n->size_layer[k] * sizeof(int32_t));
            else for (j = 0; j < n->size_layer[k]; j++) logits[j] = ((int16_t *)out)[j];
        }
    }
    load_hc(n, hl, cl);
}

static void fill_post(int16_t *post, const NNSPClass *p, int ran_nn, int stage)
{
    int i;
    post[0] = p->trigger;
    for (i = 0; i < 3; i++) post[1 + i] = p->outputs[i];
    for (i = 0; i < 8; i++) post[4 + i] = p->counts_category[i];
    post[12] = p->argmax_last;
    post[13] = p->slides;
    post[14] = (int16_t)ran_nn;
    post[15] = (int16_t)stage;
}

/* ---- one network evaluation on an explicit input and explicit LSTM state ---------------------- */
/* h/c hold the state of all lstm layers back to back (in/out); act = layers 0..L-2, logits = layer L-1 */
int ref_net_eval(int nn_id, const int16_t *input, int16_t *h, int32_t *c, int16_t *act, int32_t *logits)
{
    void *netv; const int32_t *mean, *stdR;
    static int16_t in[512], hl[1024]; static int32_t cl[1024];
    static int32_t out[512];
    if (ref_model(nn_id, &netv, &mean, &stdR)) return -1;
    NeuralNetClass *net = (NeuralNetClass *)netv;
    memcpy(in, input, net->size_layer[0] * sizeof(int16_t));
    save_hc(net, hl, cl);
    shadow_layers(net, in, h, c, act, logits);       /* per-layer taps via debug_layer, state restored */
    load_hc(net, h, c);
    NeuralNetClass_exe(net, in, out, -1);            /* the real evaluation advances the state */
    save_hc(net, h, c);
    load_hc(net, hl, cl);
    return 0;
}

/* ---- the same on a model the glue builds from an NNSPM1 container --------------------------------- */
/* The container (nnsp_b200_model_to_blob; tests/common.py make_blob) stores every table in the layout of def_nn*.c, so a
 * NeuralNetClass literal pointing straight into it is what the converter would have emitted. This is how the unmodified
 * reference evaluates SYNTHETIC layer stacks (Q-format spreads, odd widths, two LSTMs, crafted weights) for the fixtures
 * of tests/golden/make_golden_nets.py. acc32: 0 = fc_8x16 / lstm_8x16, 1 = the _acc32b twins. */
int ref_net_eval_blob(const void *blob, long long nbytes, int acc32, const int16_t *input, int16_t *h, int32_t *c,
                      int16_t *act, int32_t *logits)
{
    static NeuralNetClass net;
    static int16_t hstate[10][1024]; static int32_t cstate[10][1024];
    static int16_t in[512];
    static int32_t out[512];
    const unsigned char *p = (const unsigned char *)blob;
    int32_t nl, rec[10];
    int i;
    size_t off = 8 + 8 + 24 + 320 + 400;
    if (nbytes < (long long)off || memcmp(p, "NNSPM1\0\0", 8) != 0) return -1;
    memcpy(&nl, p + 12, 4);
    if (nl < 1 || nl > 10) return -1;
    memset(&net, 0, sizeof net);
    net.numlayers = (int8_t)nl;
    memcpy(net.size_layer, p + 16, 22);
    for (i = 0; i < nl; i++) {
        memcpy(rec, p + 360 + 40 * i, 40);
        const int is_lstm = rec[0] == 1;
        net.net_layer_type[i] = is_lstm ? lstm : fc;
        net.qbit_kernel[i] = (int8_t)rec[2]; net.qbit_input[i] = (int8_t)rec[3]; net.qbit_bias[i] = (int8_t)rec[4];
        if (i + 1 < 10) net.qbit_input[i + 1] = (int8_t)rec[9];        /* neural_nets.c:108 reads qbit_input[i+1] */
        switch (rec[1]) {
        case 0: net.activation_type[i] = relu6;   net.act_func[i] = (void *(*)(void *, int32_t *, int))&relu6_fix; break;
        case 1: net.activation_type[i] = ftanh;   net.act_func[i] = (void *(*)(void *, int32_t *, int))&tanh_fix; break;
        case 2: net.activation_type[i] = sigmoid; net.act_func[i] = (void *(*)(void *, int32_t *, int))&sigmoid_fix; break;
        default: net.activation_type[i] = linear; net.act_func[i] = (void *(*)(void *, int32_t *, int))&linear_fix; break;
        }
        if (is_lstm) net.layer_func[i] = acc32 ? (int *(*)())&lstm_8x16_acc32b : (int *(*)())&lstm_8x16;
        else         net.layer_func[i] = acc32 ? (int *(*)())&fc_8x16_acc32b : (int *(*)())&fc_8x16;
        net.pt_kernel[i] = (int8_t *)(p + off);     off += ((size_t)rec[6] + 3) & ~(size_t)3;
        net.pt_kernel_rec[i] = (int8_t *)(p + off); off += ((size_t)rec[7] + 3) & ~(size_t)3;
        net.pt_bias[i] = (int16_t *)(p + off);      off += ((size_t)rec[8] * 2 + 3) & ~(size_t)3;
        net.pt_hstate[i] = hstate[i]; net.pt_cstate[i] = cstate[i];
        if ((long long)off > nbytes) return -1;
    }
    memcpy(in, input, net.size_layer[0] * sizeof(int16_t));
    shadow_layers(&net, in, h, c, act, logits);
    load_hc(&net, h, c);
    NeuralNetClass_exe(&net, in, out, -1);
    save_hc(&net, h, c);
    return 0;
}

/* ---- one stream through NNSPClass (nn_speech.c:23-127) ----------------------------------- */
static NNSPClass g_inst;
static FeatureClass g_feat;
static int16_t g_thresh_prob, g_th_count;

int ref_nnsp_run(int nn_id, int do_reset, const int16_t *pcm, int n_frames,
                 int16_t thresh_prob, int16_t th_count, glue_result *results,
                 int32_t *tap_logmel, int16_t *tap_feat, int16_t *tap_act, int32_t *tap_logits,
                 int16_t *tap_h, int32_t *tap_c, int16_t *tap_post)
{
    void *netv; const int32_t *mean, *stdR;
    static int16_t frame[160], h0[1024]; static int32_t c0[1024];
    static FeatureClass shadow;
    int t, i;
    if (ref_model(nn_id, &netv, &mean, &stdR)) return -1;
    NeuralNetClass *net = (NeuralNetClass *)netv;
    const int as = act_stride(net), hs = h_stride(net), no = net->size_layer[net->numlayers];
    if (do_reset == 1) {            /* brand-new instance: zeroed structs, init, reset */
        g_thresh_prob = thresh_prob; g_th_count = th_count;
        memset(&g_inst, 0, sizeof g_inst); memset(&g_feat, 0, sizeof g_feat);
        NNSPClass_init(&g_inst, net, &g_feat, (char)nn_id, mean, stdR, &g_thresh_prob, &g_th_count);
        NNSPClass_reset(&g_inst);
    } else if (do_reset == 2) {     /* NNSPClass_reset of the live instance (what the controllers do) */
        NNSPClass_reset(&g_inst);
    }
    for (t = 0; t < n_frames; t++) {
        memcpy(frame, pcm + (size_t)t * 160, sizeof frame);
        const int ran = (g_inst.slides == 1);
        if (ran && (tap_act || tap_logits)) {
            shadow = g_feat;
            FeatureClass_execute(&shadow, frame);
            save_hc(net, h0, c0);
            shadow_layers(net, shadow.normFeatContext, h0, c0,
                          tap_act ? tap_act + (size_t)t * as : 0, tap_logits ? tap_logits + (size_t)t * no : 0);
        } else {
            if (tap_act) memset(tap_act + (size_t)t * as, 0, as * sizeof(int16_t));
            if (tap_logits) memset(tap_logits + (size_t)t * no, 0, no * sizeof(int32_t));
        }
        int16_t trig = NNSPClass_exec(&g_inst, frame);
        if (results) { results[t].trigger = trig; for (i = 0; i < 3; i++) results[t].outputs[i] = g_inst.outputs[i]; }
        if (tap_logmel) memcpy(tap_logmel + (size_t)t * 40, g_feat.feature, 40 * sizeof(int32_t));
        if (tap_feat) memcpy(tap_feat + (size_t)t * 40, g_feat.normFeatContext + 5 * 40, 40 * sizeof(int16_t));
        if (tap_h || tap_c) {
            save_hc(net, h0, c0);
            if (tap_h) memcpy(tap_h + (size_t)t * hs, h0, hs * sizeof(int16_t));
            if (tap_c) memcpy(tap_c + (size_t)t * hs, c0, hs * sizeof(int32_t));
        }
        if (tap_post) fill_post(tap_post + (size_t)t * 16, &g_inst, ran, nn_id);
    }
    return 0;
}

/* ---- one stream through nnCntrlClass (evb/src/nnCntrlClass.c:56-272) ----------------------- */
extern NNSPClass NNSP_INSTS[];
extern FeatureClass FEAT_INSTS[];
extern const char *intents[];
extern const char *slots[];

static int g_evt_detect;          /* set by the print hook below */
static int16_t g_evt_outputs[3];

/* oracle/ref_shim/ns_ambiqsuite_harness.h routes the controller's ns_lp_printf here: the
 * detection messages (nnCntrlClass.c:191-194,227,258) are the only place the reference
 * exposes a detection and the S2I outputs before it resets the instance. */
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
    } else if (strstr(fmt, "Detected: Hi Galaxy") || strstr(fmt, "Voice detected")) {
        g_evt_detect = 1;
    }
    va_end(ap);
}

static nnCntrlClass g_cntrl;
static NNSP_ID g_seq[8];

/* valid[t] = 1 when the per-instance taps of frame t were read from a live (not just reset)
 * instance; taps of frames on which the controller reset the instance are zero-filled. */
int ref_cascade_run(int do_reset, const int *seq, int len_seq, const int16_t *params10,
                    const int16_t *pcm, int n_frames, glue_cascade_result *results,
                    int32_t *tap_logmel, int16_t *tap_feat, int16_t *tap_h, int32_t *tap_c,
                    int16_t *tap_post, int8_t *valid)
{
    static int16_t frame[160], chunk[160], hbuf[1024]; static int32_t cbuf[1024];
    int t, i;
    if (do_reset == 2) {             /* nnCntrlClass_reset of the live controller (position kept) */
        nnCntrlClass_reset(&g_cntrl);
    } else if (do_reset) {
        for (i = 0; i < len_seq; i++) g_seq[i] = (NNSP_ID)seq[i];
        /* a brand-new controller, as after program start: the reference keeps its instances in zero-
         * initialised statics (nnCntrlClass.c:50-52) that neither _init nor _reset clears completely */
        memset(NNSP_INSTS, 0, 3 * sizeof(NNSPClass));
        memset(FEAT_INSTS, 0, 3 * sizeof(FeatureClass));
        nnCntrlClass_init(&g_cntrl, (void *)g_seq, (int8_t)len_seq);
        if (params10) memcpy(&g_cntrl.Params, params10, sizeof(ParamCntrlClass));
        nnCntrlClass_reset(&g_cntrl);
    } else if (params10) {
        /* the application changes the live controller's parameters: Params is a member of the instance and the NNSPClass
         * instances hold POINTERS into it (nnCntrlClass.c:100-123), so the next exec already sees the new values */
        memcpy(&g_cntrl.Params, params10, sizeof(ParamCntrlClass));
    }
    for (t = 0; t < n_frames; t++) {
        const int pos0 = g_cntrl.current_pos_seq, id = g_seq[pos0];
        NNSPClass *inst = &NNSP_INSTS[id];
        NeuralNetClass *net = (NeuralNetClass *)inst->pt_net;
        const int ran = (inst->slides == 1), hs_max = 128;
        memcpy(frame, pcm + (size_t)t * 160, sizeof frame);
        g_evt_detect = 0; g_evt_outputs[0] = g_evt_outputs[1] = g_evt_outputs[2] = 0;
        nnCntrlClass_exec(&g_cntrl, frame, chunk);
        const int was_reset = g_evt_detect || (g_cntrl.current_pos_seq != pos0);
        if (results) {
            results[t].stage_id = (int8_t)id;
            results[t].pos_after = g_cntrl.current_pos_seq;
            results[t].detected = (int16_t)g_evt_detect;
            for (i = 0; i < 3; i++) results[t].outputs[i] = g_evt_outputs[i];
            results[t].cnt_timeout = (id == kws_galaxy_id) ? g_cntrl.cnt_timeout_kws
                                   : (id == s2i_id)        ? g_cntrl.cnt_timeout_s2i : 0;
        }
        if (valid) valid[t] = (int8_t)!was_reset;
        if (tap_logmel) memcpy(tap_logmel + (size_t)t * 40, FEAT_INSTS[id].feature, 40 * sizeof(int32_t));
        if (tap_feat) {
            if (was_reset) memset(tap_feat + (size_t)t * 40, 0, 40 * sizeof(int16_t));
            else memcpy(tap_feat + (size_t)t * 40, FEAT_INSTS[id].normFeatContext + 5 * 40, 40 * sizeof(int16_t));
        }
        if (tap_h || tap_c) {
            memset(hbuf, 0, sizeof hbuf); memset(cbuf, 0, sizeof cbuf);
            if (!was_reset) save_hc(net, hbuf, cbuf);
            if (tap_h) memcpy(tap_h + (size_t)t * hs_max, hbuf, hs_max * sizeof(int16_t));
            if (tap_c) memcpy(tap_c + (size_t)t * hs_max, cbuf, hs_max * sizeof(int32_t));
        }
        if (tap_post) {
            if (was_reset) { memset(tap_post + (size_t)t * 16, 0, 16 * sizeof(int16_t)); tap_post[(size_t)t * 16 + 15] = (int16_t)id; }
            else fill_post(tap_post + (size_t)t * 16, inst, ran, id);
        }
    }
    return 0;
}

/* ---- many streams through ONE controller: per-stream state saved and restored around each chunk ------------ */
/* The reference is single-instance, so a process that serves several streams in turn has to
 * swap the controller's whole state in and out: the controller struct, the three NNSPClass / FeatureClass instances,
 * the PCM ring, and the LSTM state arrays of the three model tables. bench.py's CPU arms use this to keep every sampled
 * stream in the same steady state as the GPU's (which stage it is in, its time-out counters), step after step. */
#ifndef GLUE_DROPIN
extern PcmBufClass pcmbuf_inst;
extern int16_t PCM_BUFFER[];
#define GLUE_RING_SAMPLES (160 * 100)           /* PcmBufClass.c:6-7 */
#define GLUE_HC 128

typedef struct {
    nnCntrlClass cntrl;
    NNSPClass insts[3];
    FeatureClass feats[3];
    PcmBufClass ring;
    int16_t pcm[GLUE_RING_SAMPLES];
    int16_t h[3][GLUE_HC];
    int32_t c[3][GLUE_HC];
} glue_cascade_state;

long long ref_cascade_state_bytes(void) { return (long long)sizeof(glue_cascade_state); }

void ref_cascade_state_save(void *buf)
{
    glue_cascade_state *s = (glue_cascade_state *)buf;
    int id;
    s->cntrl = g_cntrl;
    memcpy(s->insts, NNSP_INSTS, sizeof s->insts);
    memcpy(s->feats, FEAT_INSTS, sizeof s->feats);
    s->ring = pcmbuf_inst;
    memcpy(s->pcm, PCM_BUFFER, sizeof s->pcm);
    for (id = 0; id < 3; id++) {
        void *netv; const int32_t *m, *sd;
        ref_model(id, &netv, &m, &sd);
        memset(s->h[id], 0, sizeof s->h[id]); memset(s->c[id], 0, sizeof s->c[id]);
        save_hc((NeuralNetClass *)netv, s->h[id], s->c[id]);
    }
}

void ref_cascade_state_load(const void *buf)
{
    const glue_cascade_state *s = (const glue_cascade_state *)buf;
    int id;
    g_cntrl = s->cntrl;
    memcpy(NNSP_INSTS, s->insts, sizeof s->insts);
    memcpy(FEAT_INSTS, s->feats, sizeof s->feats);
    pcmbuf_inst = s->ring;
    memcpy(PCM_BUFFER, s->pcm, sizeof s->pcm);
    for (id = 0; id < 3; id++) {
        void *netv; const int32_t *m, *sd;
        ref_model(id, &netv, &m, &sd);
        load_hc((NeuralNetClass *)netv, s->h[id], s->c[id]);
    }
}

/* n_streams streams, n_frames each, through nnCntrlClass_exec; states: n_streams blobs of ref_cascade_state_bytes()
 * (fresh != 0: every stream starts from a new controller). stage_frames[3] (may be null) counts frames per NNSP id. */
int ref_cascade_batch(int fresh, const int *seq, int len_seq, const int16_t *params10, int n_streams,
                      const int16_t *pcm, long long stream_stride, int n_frames, void *states,
                      glue_cascade_result *results, long long *stage_frames)
{
    int s;
    const long long sb = ref_cascade_state_bytes();
    for (s = 0; s < n_streams; s++) {
        char *st = (char *)states + (size_t)s * sb;
        const int16_t *x = pcm + (size_t)s * stream_stride;
        int rc;
        if (fresh) rc = ref_cascade_run(1, seq, len_seq, params10, x, 0, 0, 0, 0, 0, 0, 0, 0);
        else { ref_cascade_state_load(st); rc = 0; }
        if (rc) return rc;
        if (results || stage_frames) {
            static glue_cascade_result tmp[4096];
            int t0;
            for (t0 = 0; t0 < n_frames; t0 += 4096) {
                const int n = (n_frames - t0) < 4096 ? (n_frames - t0) : 4096;
                int t;
                glue_cascade_result *r = results ? results + (size_t)s * n_frames + t0 : tmp;
                ref_cascade_run(0, seq, len_seq, params10, x + (size_t)t0 * 160, n, r, 0, 0, 0, 0, 0, 0);
                if (stage_frames) for (t = 0; t < n; t++) stage_frames[r[t].stage_id]++;
            }
        } else {
            ref_cascade_run(0, seq, len_seq, params10, x, n_frames, 0, 0, 0, 0, 0, 0, 0);
        }
        ref_cascade_state_save(st);
    }
    return 0;
}
#endif
