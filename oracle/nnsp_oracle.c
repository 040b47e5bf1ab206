/* oracle/nnsp_oracle.c -- TEST INFRASTRUCTURE (see nnsp_oracle.h): CPU restatement of the
 * ns-nnsp streaming hot path, function by function, citing the reference lines it follows.
 * Compile with -fwrapv (the reference's 32-bit accumulator mode relies on wrap-around).
 * Parity pinned against oracle/_ref (the compiled reference) by tests/test_oracle_vs_ref.py. */
#include "nnsp_oracle.h"
#include "nnsp_model.h"
#include "nnsp_tables.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define I32_MAX ((int64_t)0x7fffffff)
#define I32_MIN (-(int64_t)0x80000000LL)
static int64_t clamp64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int32_t sat32(int64_t v) { return (int32_t)clamp64(v, I32_MIN, I32_MAX); }

typedef struct { int32_t re, im; } cpx32;

/* ======================================================================================== */
/* Front end                                                                                 */
/* ======================================================================================== */
static cpx32 unpack_tw(int32_t packed)  /* COMPLEX16 overlay: low half real, high half imag (complex.h:10-14) */
{
    cpx32 t = { (int16_t)(packed & 0xffff), (int16_t)((uint32_t)packed >> 16) };
    return t;
}

/* complex32_complex16_elmtprod, complex.c:54-72 */
static cpx32 cmul_q15(cpx32 a, cpx32 w)
{
    int64_t re = (int64_t)a.re * w.re - (int64_t)a.im * w.im;
    int64_t im = (int64_t)a.re * w.im + (int64_t)a.im * w.re;
    cpx32 r = { sat32(re >> 15), sat32(im >> 15) };
    return r;
}

/* fft(), fft.c:128-221, for the only size the front end uses (256 complex points, 4 radix-4
 * DIF stages). The 4x4 matrix product of complex32_affine (complex.c:14-52) with M4_
 * (fft.c:12-15) is written out; sums are formed in 64 bits and saturated like the original. */
static void fft256(cpx32 *x, cpx32 *out, const nnsp_tables *T)
{
    int Nf = 256, Ng = 1, S = 1;
    for (int s = 0; s < 4; s++) {
        const int q = Nf >> 2;
        for (int g = 0; g < Ng; g++) {
            int k = 0;
            for (int m = 0; m < q; m++) {
                const int i0 = g * Nf + m;
                const cpx32 a = x[i0], c = x[i0 + q], b = x[i0 + 2 * q], d = x[i0 + 3 * q];   /* fft.c:171-178 */
                cpx32 o[4];
                o[0].re = sat32((int64_t)a.re + b.re + c.re + d.re);
                o[0].im = sat32((int64_t)a.im + b.im + c.im + d.im);
                o[1].re = sat32((int64_t)a.re + b.re - c.re - d.re);
                o[1].im = sat32((int64_t)a.im + b.im - c.im - d.im);
                o[2].re = sat32((int64_t)a.re - b.re + c.im - d.im);
                o[2].im = sat32((int64_t)a.im - b.im - c.re + d.re);
                o[3].re = sat32((int64_t)a.re - b.re - c.im + d.im);
                o[3].im = sat32((int64_t)a.im - b.im + c.re - d.re);
                for (int n = 0; n < 4; n++) o[n] = cmul_q15(o[n], unpack_tw(T->fft_tw[4 * k + n]));  /* fft.c:182 */
                k += S;
                x[i0] = o[0]; x[i0 + q] = o[1]; x[i0 + 2 * q] = o[2]; x[i0 + 3 * q] = o[3];           /* fft.c:185-192 */
            }
        }
        Nf >>= 2; Ng <<= 2; S <<= 2;
    }
    for (int m = 0; m < 256; m++) out[m] = x[T->bitrev[m]];                                            /* fft.c:217-220 */
}

/* rfft(512,...), fft.c:27-126 */
static void rfft512(const int32_t *in, cpx32 *X /*[257]*/, const nnsp_tables *T)
{
    cpx32 cin[256], Z[256], Xe[256], Xo[256];
    for (int i = 0; i < 256; i++) { cin[i].re = in[2 * i]; cin[i].im = in[2 * i + 1]; }
    fft256(cin, Z, T);
    for (int i = 0; i < 256; i++) {
        const cpx32 zi = Z[i], zr = Z[(256 - i) & 255];
        Xe[i].re = (zi.re + zr.re) >> 1;            /* fft.c:68-69, 91-92 */
        Xe[i].im = (zi.im - zr.im) >> 1;
        Xo[i].re = (zi.im + zr.im) >> 1;            /* fft.c:71, 94 */
        Xo[i].im = (-(zi.re - zr.re)) >> 1;         /* fft.c:72, 95: unary minus binds before >> */
    }
    for (int i = 0; i < 256; i++) {
        const cpx32 p = cmul_q15(Xo[i], unpack_tw(T->rfft_tw[i]));    /* fft.c:107-115 */
        X[i].re = Xe[i].re + p.re;                                     /* fft.c:120 */
        X[i].im = Xe[i].im + p.im;
    }
    X[256].re = Xe[0].re - Xo[0].re;                                   /* fft.c:123-124 */
    X[256].im = Xe[0].im - Xo[0].im;
}

/* my_log10 + norm_oneTwo + log10_vec(.., bit_frac_in = 15), fixlog10.c:9-61 */
static int32_t log10_q15(int32_t x, const nnsp_tables *T)
{
    if (x == 0) x = 1;
    int sh = 0;
    for (int i = 0; i < 31; i++)
        if (x & ((int32_t)1 << (30 - i))) { sh = -(30 - i - 15); break; }
    const int32_t y = (sh >= 0) ? (int32_t)((uint32_t)x << sh) : x >> -sh;
    sh = -sh;
    const int32_t kx = (y - 32768) >> 8, dx = (y - 32768) - (kx << 8);
    int32_t t = (int32_t)T->log_lut[kx << 1] + (((int32_t)T->log_lut[1 + (kx << 1)] * dx) >> 15);
    t = (int32_t)(((int64_t)t * 0x3796) >> 15);
    return t + 0x2688 * (int32_t)(int8_t)sh;
}

static void window_frame(const int16_t *buf480, int32_t *fft_in, const nnsp_tables *T)
{
    for (int i = 0; i < 480; i++) fft_in[i] = ((int32_t)T->stft_win[i] * (int32_t)buf480[i]) >> 15;   /* spectrogram_module.c:62-66 */
    for (int i = 480; i < 512; i++) fft_in[i] = 0;                                                    /* :68-71 */
}

static void spectrum_to_logmel(const cpx32 *X, int32_t *pspec, int32_t *mel, int32_t *logmel, const nnsp_tables *T)
{
    for (int i = 0; i < 257; i++) {                         /* spec2pspec, spectrogram_module.c:33-45 (truncating cast) */
        const int64_t p = (int64_t)X[i].re * X[i].re + (int64_t)X[i].im * X[i].im;
        pspec[i] = (int32_t)(p >> 15);
    }
    const int16_t *c = T->mel;                              /* melSpecProc, melSpecProc.c:6-27 */
    for (int b = 0; b < 40; b++) {
        const int s = *c++, e = *c++;
        int64_t mac = 0;
        for (int j = s; j <= e; j++) mac += (int64_t)(*c++) * (int64_t)pspec[j];
        mel[b] = sat32(mac >> 15);
    }
    for (int b = 0; b < 40; b++) logmel[b] = log10_q15(mel[b], T);
}

void nnsp_oracle_feature_stages(const int16_t *win480, int32_t *fft_in, int32_t *spec,
                                int32_t *pspec, int32_t *mel, int32_t *logmel)
{
    const nnsp_tables *T = nnsp_tables_get();
    int32_t fi[512], ps[257], me[40], lm[40];
    cpx32 X[257];
    window_frame(win480, fi, T);
    rfft512(fi, X, T);
    spectrum_to_logmel(X, ps, me, lm, T);
    if (fft_in) memcpy(fft_in, fi, sizeof fi);
    if (spec) memcpy(spec, X, sizeof X);
    if (pspec) memcpy(pspec, ps, sizeof ps);
    if (mel) memcpy(mel, me, sizeof me);
    if (logmel) memcpy(logmel, lm, sizeof lm);
}

/* FeatureClass state (feature_module.h:7-18, spectrogram_module.h:9-16) */
typedef struct {
    int16_t buf[480];
    int32_t feature[40];
    int16_t ctx[6 * 40];
} o_feat;

/* (x - mean) * stdR >> (30 - qbit_output), saturate to int16: feature_module.c:34-37, 69-72 */
static int16_t standardise(int32_t v, int32_t mean, int32_t stdR, int qout)
{
    int64_t t = (int64_t)v - (int64_t)mean;
    t = (t * (int64_t)stdR) >> (30 - qout);
    return (int16_t)clamp64(t, -32768, 32767);
}

static void feat_set_default(o_feat *f, const nnsp_b200_model *m)      /* feature_module.c:26-45 */
{
    memset(f->buf, 0, sizeof f->buf);                                  /* stftModule_setDefault, spectrogram_module.c:25-31 */
    for (int i = 0; i < 40; i++) {
        const int16_t v = standardise(-147963, m->mean[i], m->stdR[i], m->layer[0].qi);
        for (int j = 0; j < 5; j++) f->ctx[i + j * 40] = v;
    }
}

static void feat_execute(o_feat *f, const nnsp_b200_model *m, const int16_t *pcm160, const nnsp_tables *T)  /* feature_module.c:47-75 */
{
    int32_t fi[512], ps[257], me[40];
    cpx32 X[257];
    memmove(f->ctx, f->ctx + 40, 200 * sizeof(int16_t));               /* :54-57 */
    memmove(f->buf, f->buf + 160, 320 * sizeof(int16_t));              /* spectrogram_module.c:55-60 */
    memcpy(f->buf + 320, pcm160, 160 * sizeof(int16_t));
    window_frame(f->buf, fi, T);
    rfft512(fi, X, T);
    spectrum_to_logmel(X, ps, me, f->feature, T);
    for (int i = 0; i < 40; i++)
        f->ctx[200 + i] = standardise(f->feature[i], m->mean[i], m->stdR[i], m->layer[0].qi);
}

/* ======================================================================================== */
/* Network                                                                                    */
/* ======================================================================================== */
/* tanh_fix, activation.c:31-69. x == INT32_MIN makes the reference index its LUT out of
 * bounds (undefined); this restatement and the CUDA engine both return -0x7fff there. */
static int16_t tanh_q15(int32_t x, const nnsp_tables *T)
{
    const int neg = x < 0;
    int32_t xi = neg ? (int32_t)(0u - (uint32_t)x) : x;
    int16_t y;
    if (xi < 0 || xi >= (5 << 15)) y = 0x7fff;
    else {
        int32_t kx = (xi - 512) >> 10;
        if (kx < 0) kx = 0;
        int32_t dx = xi - 512 - (kx << 10);
        dx = (int32_t)T->tanh_lut[kx << 1] + ((dx * (int32_t)T->tanh_lut[(kx << 1) + 1]) >> 15);
        y = (int16_t)(dx > 0 ? dx : 0);
    }
    return neg ? (int16_t)-y : y;
}
static int16_t sigmoid_q15(int32_t x, const nnsp_tables *T)     /* activation.c:72-86 */
{
    int16_t y = tanh_q15(x >> 1, T);
    y = (int16_t)(y >> 1);
    return (int16_t)(y + 16384);
}
static int16_t relu6_q12(int32_t x)                              /* activation.c:6-17 */
{
    int32_t v = x >> 3;
    if (v > (6 << 12)) v = 6 << 12;
    if (v < 0) v = 0;
    return (int16_t)v;
}

/* shift_64b, affine.c:565-591 / shift_32b, affine_acc32b.c:566-592 */
static int64_t shift64(int64_t x, int sh)
{
    if (sh == 0) return x;
    if (sh < 0) return x >> -sh;
    int64_t M = (int64_t)1 << (63 - sh), m = -M;
    M -= 1;
    x = clamp64(x, m, M);
    return (int64_t)((uint64_t)x << sh);
}
static int32_t shift32(int32_t x, int sh)
{
    if (sh == 0) return x;
    if (sh < 0) return x >> -sh;
    int32_t M = (int32_t)((uint32_t)1 << (31 - sh)), m = (int32_t)(0u - (uint32_t)M);
    M -= 1;
    if (x < m) x = m;
    if (x > M) x = M;
    return (int32_t)((uint32_t)x << sh);
}

/* One output row through the ARM-variant affine_Krows_8x16 (affine.c:12-259) or its acc32b
 * twin (affine_acc32b.c:12-260). `acc` is the running accumulator (64-bit value; in acc32
 * mode only its low 32 bits are meaningful and every step wraps). Returns the new accumulator;
 * when is_out, *pre receives the 32-bit pre-activation.
 * Quirk Q1: the "align accumulator to the bias Q-format" shift at
 * affine.c:186-187 acts on memory that is overwritten at :219-240, so it is a no-op here. */
static int64_t affine_row(int64_t acc, const int8_t *w, const int16_t *x, int cols, const int16_t *bias,
                          int qk, int qb, int qi, int acc32, int is_out, int32_t *pre)
{
    const int qs = bias ? ((qi + qk) > 15 ? (qi + qk) : 15) : (qi + qk);       /* affine.c:69-72 */
    if (!acc32) {
        for (int c = 0; c < cols; c++) acc += (int64_t)w[c] * (int64_t)x[c];
        if (bias) {
            const int sh = qs - qb;                                             /* affine.c:190-199 */
            acc += (sh >= 0) ? (int64_t)((uint64_t)(int64_t)*bias << sh) : ((int64_t)*bias >> -sh);
        }
        if (is_out) {
            const int64_t o = shift64(acc, (int8_t)(15 - qs));                  /* affine.c:244-248 */
            *pre = sat32(o);
            return o;
        }
        return acc;
    }
    uint32_t a = (uint32_t)acc;
    for (int c = 0; c < cols; c++) a += (uint32_t)((int32_t)w[c] * (int32_t)x[c]);
    if (bias) {
        const int sh = qs - qb;
        a += (sh >= 0) ? ((uint32_t)(int32_t)*bias << sh) : (uint32_t)((int32_t)*bias >> -sh);
    }
    if (is_out) {
        const int32_t o = shift32((int32_t)a, (int8_t)(15 - qs));               /* affine_acc32b.c:245-249: no saturation */
        *pre = o;
        return (int64_t)o;
    }
    return (int64_t)(int32_t)a;
}

static void apply_act(int act, int32_t pre, int16_t *y16, int32_t *y32, const nnsp_tables *T)
{
    switch (act) {
    case NNSP_ACT_TANH: *y16 = tanh_q15(pre, T); break;
    case NNSP_ACT_SIGMOID: *y16 = sigmoid_q15(pre, T); break;
    case NNSP_ACT_RELU6: *y16 = relu6_q12(pre); break;
    default: *y32 = pre; break;
    }
}

/* fc_8x16, affine.c:409-490 */
static void fc_layer(const nnsp_layer *L, const int16_t *x, int16_t *y16, int32_t *y32, const nnsp_tables *T)
{
    for (int r = 0; r < L->rows; r++) {
        int32_t pre;
        affine_row(0, L->w + (size_t)r * L->cols, x, L->cols, &L->bias[r], L->qk, L->qb, L->qi, L->acc32, 1, &pre);
        apply_act(L->act, pre, &y16[r], &y32[r], T);
    }
}

/* rc_Krows_8x16, affine.c:348-407: input half (no bias, not output), re-scale from the input's
 * Q-format to the recurrent input's, recurrent half + bias, output */
static int32_t lstm_gate_pre(const nnsp_layer *L, int row, const int16_t *x, const int16_t *h)
{
    int32_t pre = 0;
    int64_t acc = affine_row(0, L->w + (size_t)row * L->cols, x, L->cols, NULL, L->qk, L->qb, L->qi, L->acc32, 0, &pre);
    const int sh = (int8_t)(L->qi_next - L->qi);                                /* affine.c:371,384 */
    acc = L->acc32 ? (int64_t)shift32((int32_t)acc, sh) : shift64(acc, sh);
    affine_row(acc, L->wrec + (size_t)row * L->rows, h, L->rows, &L->bias[row], L->qk, L->qb, L->qi_next, L->acc32, 1, &pre);
    return pre;
}

/* lstm_8x16, lstm.c:15-214: gates i, j(g), f, o; c = (i*j + f*c) >> 15 saturated to int32;
 * out = sat16((tanh(c) * o) >> 15); every gate sees the OLD h; h is committed at the end */
static void lstm_layer(const nnsp_layer *L, const int16_t *x, int16_t *h, int32_t *c, int16_t *y, const nnsp_tables *T)
{
    const int H = L->rows;
    for (int u = 0; u < H; u++) {
        const int16_t gi = sigmoid_q15(lstm_gate_pre(L, 0 * H + u, x, h), T);
        const int16_t gj = tanh_q15(lstm_gate_pre(L, 1 * H + u, x, h), T);
        const int16_t gf = sigmoid_q15(lstm_gate_pre(L, 2 * H + u, x, h), T);
        const int16_t go = sigmoid_q15(lstm_gate_pre(L, 3 * H + u, x, h), T);
        const int64_t t = ((int64_t)gi * (int64_t)gj + (int64_t)gf * (int64_t)c[u]) >> 15;    /* lstm.c:108-109 */
        c[u] = sat32(t);
        int32_t o = ((int32_t)tanh_q15(c[u], T) * (int32_t)go) >> 15;                         /* lstm.c:111-115 */
        if (o > 32767) o = 32767;
        if (o < -32768) o = -32768;
        y[u] = (int16_t)o;
    }
    memcpy(h, y, (size_t)H * sizeof(int16_t));                                                /* lstm.c:205-206 */
}

static int model_act_stride(const nnsp_b200_model *m) { int s = 0; for (int i = 1; i < m->numlayers; i++) s += m->size_layer[i]; return s; }
static int model_h_stride(const nnsp_b200_model *m) { int s = 0; for (int i = 0; i < m->numlayers; i++) if (m->layer[i].type == NNSP_LAYER_LSTM) s += m->size_layer[i + 1]; return s; }

/* NeuralNetClass_exe with debug_layer = -1, neural_nets.c:44-168. h/c hold the state of all
 * lstm layers back to back. act receives layers 0..L-2, logits the last layer (int32). */
#define O_MAXW 4096
static void net_exec(const nnsp_b200_model *m, const int16_t *input, int16_t *h, int32_t *c,
                     int16_t *act, int32_t *logits, const nnsp_tables *T)
{
    int16_t a[O_MAXW], b[O_MAXW];
    int32_t o32[O_MAXW];
    const int16_t *in = input;
    int16_t *out = a;
    int ho = 0, ao = 0;
    for (int i = 0; i < m->numlayers; i++) {
        const nnsp_layer *L = &m->layer[i];
        if (L->type == NNSP_LAYER_LSTM) { lstm_layer(L, in, h + ho, c + ho, out, T); ho += L->rows; }
        else fc_layer(L, in, out, o32, T);
        if (i < m->numlayers - 1) {
            if (act) memcpy(act + ao, out, (size_t)L->rows * sizeof(int16_t));
            ao += L->rows;
        } else if (logits) {
            const int is32 = (L->type == NNSP_LAYER_FC && L->act == NNSP_ACT_LINEAR);   /* neural_nets.c:152-167 */
            for (int j = 0; j < L->rows; j++) logits[j] = is32 ? o32[j] : (int32_t)out[j];
        }
        in = out;
        out = (out == a) ? b : a;
    }
}

int nnsp_oracle_net_eval(const nnsp_b200_model *m, const int16_t *input, int16_t *h, int32_t *c,
                         int16_t *act, int32_t *logits)
{
    const nnsp_tables *T = nnsp_tables_get();
    if (!T || !m) return -1;
    net_exec(m, input, h, c, act, logits, T);
    return 0;
}

/* ======================================================================================== */
/* NNSPClass                                                                                  */
/* ======================================================================================== */
struct nnsp_oracle_stream {
    o_feat  feat;
    int16_t h[O_MAXW];
    int32_t c[O_MAXW];
    int8_t  slides;
    int16_t trigger, counts[8], outputs[3], argmax_last;
};

nnsp_oracle_stream *nnsp_oracle_stream_new(void) { return (nnsp_oracle_stream *)calloc(1, sizeof(nnsp_oracle_stream)); }
void nnsp_oracle_stream_free(nnsp_oracle_stream *s) { free(s); }

static void stream_reset(nnsp_oracle_stream *s, const nnsp_b200_model *m)       /* NNSPClass_reset, nn_speech.c:57-72 */
{
    feat_set_default(&s->feat, m);
    memset(s->h, 0, sizeof s->h);                                               /* NeuralNetClass_setDefault, neural_nets.c:27-42 */
    memset(s->c, 0, sizeof s->c);
    s->slides = 1;
    s->trigger = 0;
    for (int i = 0; i < 7; i++) s->counts[i] = 0;                               /* only DIM_INTENTS of the 8 (nn_speech.c:64-65) */
    for (int i = 0; i < 3; i++) s->outputs[i] = 0;
    s->argmax_last = 0;
}

static int16_t argmax_last_wins(const int32_t *v, int n)                        /* my_argmax, nn_speech.c:130-144 (>= : ties go to the last) */
{
    int16_t im = 0;
    int32_t mx = v[0];
    for (int i = 1; i < n; i++) if (v[i] >= mx) { mx = v[i]; im = (int16_t)i; }
    return im;
}

static void post_s2i(nnsp_oracle_stream *s, const int32_t *est, int16_t th_count)   /* nn_speech.c:146-189 */
{
    s->trigger = 0;
    for (int i = 0; i < 3; i++) s->outputs[i] = 0;
    const int16_t ai = argmax_last_wins(est, 7);
    if (s->argmax_last == 0 || s->argmax_last == ai) {
        if (ai != 0) {
            s->counts[ai] = (int16_t)(s->counts[ai] + 1);
            if (s->counts[ai] > th_count) {
                s->trigger = 1;
                s->outputs[0] = ai;
                s->outputs[1] = argmax_last_wins(est + 7, 17);
                s->outputs[2] = argmax_last_wins(est + 7 + 17, 17);
            }
        }
    } else {
        for (int i = 0; i < 7; i++) s->counts[i] = 0;
    }
    s->argmax_last = ai;
}

static int32_t pwr2_q15(int32_t in)                                             /* ceiling + compute_pwr2, nn_speech.c:229-258 */
{
    int32_t ce = (int32_t)((uint32_t)(in >> 15) << 15);
    if (ce - in != 0) ce = (int32_t)((uint32_t)ce + 32768u);
    in = (int32_t)((uint32_t)in - (uint32_t)ce);
    const int32_t shift = ce >> 15;
    if (shift <= -15) return 0;
    const int32_t t = (int32_t)(((uint32_t)in << 1) + 32768u);
    int32_t o = 0x1fd7 + ((int32_t)((uint32_t)t * (uint32_t)0x057a) >> 15);
    o = 0x5a82 + ((int32_t)((uint32_t)t * (uint32_t)o) >> 15);
    return (shift < 0) ? o >> -shift : (int32_t)((uint32_t)o << shift);
}

static void post_binary(nnsp_oracle_stream *s, const int32_t *logits, int16_t thresh_prob, int16_t th_count)  /* nn_speech.c:191-227 */
{
    int32_t est[2];
    const int32_t mx = logits[0] > logits[1] ? logits[0] : logits[1];
    for (int i = 0; i < 2; i++) {
        const int32_t val = (int32_t)((uint32_t)logits[i] - (uint32_t)mx);
        const int64_t ref = clamp64(((int64_t)val * 0xB8AA) >> 15, I32_MIN, I32_MAX);
        est[i] = pwr2_q15((int32_t)ref);
    }
    const int32_t den = (int32_t)((uint32_t)est[0] + (uint32_t)est[1]);
    const int32_t thresh = 32768 - (int)thresh_prob;
    const int32_t tmp = (int32_t)(((int64_t)thresh * (int64_t)den) >> 15);
    if (est[0] <= tmp) s->counts[0] = (int16_t)(s->counts[0] + 1);
    else s->counts[0] = 0;
    s->trigger = (s->counts[0] >= th_count) ? 1 : 0;
}

/* NNSPClass_exec, nn_speech.c:74-127. Returns whether the network ran. */
static int stream_exec(nnsp_oracle_stream *s, const nnsp_b200_model *m, const int16_t *pcm160,
                       int16_t thresh_prob, int16_t th_count, int16_t *act, int32_t *logits_out,
                       const nnsp_tables *T)
{
    int32_t logits[O_MAXW];
    feat_execute(&s->feat, m, pcm160, T);
    const int ran = (s->slides == 1);
    if (ran) {
        net_exec(m, s->feat.ctx, s->h, s->c, act, logits, T);
        if (logits_out) memcpy(logits_out, logits, (size_t)m->size_layer[m->numlayers] * sizeof(int32_t));
        if (m->nn_id == NNSP_B200_ID_S2I) post_s2i(s, logits, th_count);
        else if (m->nn_id == NNSP_B200_ID_VAD || m->nn_id == NNSP_B200_ID_KWS) post_binary(s, logits, thresh_prob, th_count);
    }
    s->slides = (int8_t)((s->slides + 1) % 2);
    return ran;
}

static void fill_post(int16_t *post, const nnsp_oracle_stream *s, int ran, int stage)
{
    post[0] = s->trigger;
    for (int i = 0; i < 3; i++) post[1 + i] = s->outputs[i];
    for (int i = 0; i < 8; i++) post[4 + i] = s->counts[i];
    post[12] = s->argmax_last;
    post[13] = s->slides;
    post[14] = (int16_t)ran;
    post[15] = (int16_t)stage;
}

int nnsp_oracle_nnsp_run(const nnsp_b200_model *m, nnsp_oracle_stream *s, int do_reset,
                         const int16_t *pcm, int n_frames, int16_t thresh_prob, int16_t th_count,
                         nnsp_b200_result *results, int32_t *tap_logmel, int16_t *tap_feat,
                         int16_t *tap_act, int32_t *tap_logits, int16_t *tap_h, int32_t *tap_c,
                         int16_t *tap_post)
{
    const nnsp_tables *T = nnsp_tables_get();
    if (!T || !m || !s) return -1;
    const int as = model_act_stride(m), hs = model_h_stride(m), no = m->size_layer[m->numlayers];
    if (do_reset == 1) { memset(s, 0, sizeof *s); stream_reset(s, m); }   /* brand-new instance */
    else if (do_reset == 2) stream_reset(s, m);                           /* NNSPClass_reset of a live instance: context row 5 survives */
    for (int t = 0; t < n_frames; t++) {
        int16_t act[O_MAXW];
        int32_t logits[O_MAXW];
        memset(act, 0, (size_t)as * sizeof(int16_t));
        memset(logits, 0, (size_t)no * sizeof(int32_t));
        const int ran = stream_exec(s, m, pcm + (size_t)t * 160, thresh_prob, th_count, act, logits, T);
        if (results) { results[t].trigger = s->trigger; memcpy(results[t].outputs, s->outputs, sizeof s->outputs); }
        if (tap_logmel) memcpy(tap_logmel + (size_t)t * 40, s->feat.feature, 40 * sizeof(int32_t));
        if (tap_feat) memcpy(tap_feat + (size_t)t * 40, s->feat.ctx + 200, 40 * sizeof(int16_t));
        if (tap_act) memcpy(tap_act + (size_t)t * as, act, (size_t)as * sizeof(int16_t));
        if (tap_logits) memcpy(tap_logits + (size_t)t * no, logits, (size_t)no * sizeof(int32_t));
        if (tap_h) memcpy(tap_h + (size_t)t * hs, s->h, (size_t)hs * sizeof(int16_t));
        if (tap_c) memcpy(tap_c + (size_t)t * hs, s->c, (size_t)hs * sizeof(int32_t));
        if (tap_post) fill_post(tap_post + (size_t)t * 16, s, ran, m->nn_id);
    }
    return 0;
}

/* ======================================================================================== */
/* nnCntrlClass + PcmBufClass                                                                 */
/* ======================================================================================== */
#define RING_FRAMES 100                                        /* NUM_FRS_VBUF, PcmBufClass.c:6 */
struct nnsp_oracle_cascade {
    nnsp_oracle_stream inst[3];                                /* NNSP_INSTS[id], nnCntrlClass.c:50 */
    int16_t ring[RING_FRAMES * 160];
    int16_t idx_set, idx_latest;
    int     seq[8], len_seq, pos;
    nnsp_b200_cascade_params P;
    uint16_t cnt_kws, cnt_s2i;
};

nnsp_oracle_cascade *nnsp_oracle_cascade_new(void) { return (nnsp_oracle_cascade *)calloc(1, sizeof(nnsp_oracle_cascade)); }
void nnsp_oracle_cascade_free(nnsp_oracle_cascade *c) { free(c); }

void nnsp_oracle_default_params(nnsp_b200_cascade_params *p)   /* ParamsNNCntrl.h:8-21 */
{
    p->thresh_prob_vad = 32767 >> 1; p->thresh_cnts_vad = 4;
    p->frs_vbufBk_s2i = 80; p->thresh_timeout_s2i = 1000; p->thresh_prob_s2i = 32767 >> 1; p->thresh_cnts_s2i = 4;
    p->frs_vbufBk_kws = 80; p->thresh_timeout_kws = 1000; p->thresh_prob_kws = 32767 >> 1; p->thresh_cnts_kws = 4;
}

static void ring_get(const nnsp_oracle_cascade *c, int lookback, int16_t *out)          /* PcmBufClass_getData, PcmBufClass.c:52-85, 1 frame */
{
    int16_t start = (int16_t)((c->idx_latest - lookback) % RING_FRAMES);
    if (start < 0) start = (int16_t)(start + RING_FRAMES);
    memcpy(out, c->ring + (size_t)start * 160, 160 * sizeof(int16_t));
}

static int cascade_step(nnsp_oracle_cascade *c, const nnsp_b200_model *const models[3], const int16_t *frame,
                        nnsp_b200_cascade_result *res, const nnsp_tables *T, int *ran_out, int *reset_out)
{
    const int id = c->seq[c->pos];
    nnsp_oracle_stream *s = &c->inst[id];
    const nnsp_b200_model *m = models[id];
    int16_t chunk[160];
    int next_pos = c->pos, detected = 0, did_reset = 0;
    uint16_t cnt = 0;
    int16_t outs[3] = { 0, 0, 0 };

    memcpy(c->ring + (size_t)c->idx_set * 160, frame, 160 * sizeof(int16_t));           /* PcmBufClass_setData, PcmBufClass.c:30-50 */
    c->idx_latest = c->idx_set;
    c->idx_set = (int16_t)((c->idx_set + 1) % RING_FRAMES);

    if (id == NNSP_B200_ID_S2I) {                                                        /* nnCntrlClass.c:174-207 */
        ring_get(c, c->P.frs_vbufBk_s2i, chunk);
        *ran_out = stream_exec(s, m, chunk, c->P.thresh_prob_s2i, c->P.thresh_cnts_s2i, NULL, NULL, T);
        detected = s->trigger;
        memcpy(outs, s->outputs, sizeof outs);
        c->cnt_s2i = (uint16_t)((c->cnt_s2i + 1) % c->P.thresh_timeout_s2i);
        if (detected || c->cnt_s2i == (c->P.thresh_timeout_s2i - 1)) {
            next_pos = (c->pos + 1) % c->len_seq;
            if (detected || id != c->seq[next_pos]) { c->cnt_s2i = 0; stream_reset(s, m); did_reset = 1; }
        }
        cnt = c->cnt_s2i;
    } else if (id == NNSP_B200_ID_KWS) {                                                 /* nnCntrlClass.c:209-243 */
        ring_get(c, c->P.frs_vbufBk_kws, chunk);
        *ran_out = stream_exec(s, m, chunk, c->P.thresh_prob_kws, c->P.thresh_cnts_kws, NULL, NULL, T);
        detected = s->trigger;
        memcpy(outs, s->outputs, sizeof outs);
        c->cnt_kws = (uint16_t)((c->cnt_kws + 1) % c->P.thresh_timeout_kws);
        if (detected || c->cnt_kws == (c->P.thresh_timeout_kws - 1)) {
            if (detected) next_pos = (c->pos + 1) % c->len_seq;
            else { next_pos = (c->pos - 1) % c->len_seq; if (next_pos < 0) next_pos += c->len_seq; }
            if (detected || id != c->seq[next_pos]) { c->cnt_kws = 0; stream_reset(s, m); did_reset = 1; }
        }
        cnt = c->cnt_kws;
    } else {                                                                             /* vad, nnCntrlClass.c:245-268 */
        ring_get(c, 0, chunk);
        *ran_out = stream_exec(s, m, chunk, c->P.thresh_prob_vad, c->P.thresh_cnts_vad, NULL, NULL, T);
        detected = s->trigger;
        memcpy(outs, s->outputs, sizeof outs);
        if (detected) {
            next_pos = (c->pos + 1) % c->len_seq;
            stream_reset(s, m);
            did_reset = 1;
        }
    }
    c->pos = next_pos;
    if (res) {
        res->stage_id = (int8_t)id;
        res->pos_after = (int8_t)c->pos;
        res->detected = (int16_t)detected;
        memcpy(res->outputs, outs, sizeof outs);
        res->cnt_timeout = cnt;
    }
    *reset_out = did_reset;
    return id;
}

static void cascade_reset(nnsp_oracle_cascade *c, const nnsp_b200_model *const models[3], const int *seq, int len_seq,
                          const nnsp_b200_cascade_params *params)
{
    memset(c, 0, sizeof *c);
    for (int i = 0; i < len_seq && i < 8; i++) c->seq[i] = seq[i];
    c->len_seq = len_seq;
    if (params) c->P = *params; else nnsp_oracle_default_params(&c->P);
    c->pos = 0;                                                     /* nnCntrlClass_init, nnCntrlClass.c:125 */
    c->cnt_kws = c->cnt_s2i = 0;                                    /* nnCntrlClass_reset, :140-141 */
    for (int i = 0; i < 3; i++) if (models[i]) stream_reset(&c->inst[i], models[i]);
    c->idx_set = 0;                                                 /* PcmBufClass_reset, PcmBufClass.c:19-28 */
    c->idx_latest = RING_FRAMES - 1;
}

int nnsp_oracle_cascade_run(const nnsp_b200_model *const models[3], nnsp_oracle_cascade *c,
                            int do_reset, const int *seq, int len_seq,
                            const nnsp_b200_cascade_params *params, const int16_t *pcm,
                            int n_frames, nnsp_b200_cascade_result *results, int32_t *tap_logmel,
                            int16_t *tap_feat, int16_t *tap_h, int32_t *tap_c, int16_t *tap_post,
                            int8_t *valid)
{
    const nnsp_tables *T = nnsp_tables_get();
    if (!T || !c) return -1;
    if (do_reset == 2) {            /* nnCntrlClass_reset of a live controller, nnCntrlClass.c:130-150: position kept */
        c->cnt_kws = c->cnt_s2i = 0;
        for (int i = 0; i < 3; i++) if (models[i]) stream_reset(&c->inst[i], models[i]);
        memset(c->ring, 0, sizeof c->ring);
        c->idx_set = 0;
        c->idx_latest = RING_FRAMES - 1;
    } else if (do_reset) cascade_reset(c, models, seq, len_seq, params);
    else if (params) c->P = *params;    /* a live controller whose Params the application rewrote: the instances read the
                                         * thresholds through pointers into it (nnCntrlClass.c:100-123), the time-outs directly (:185, :219) */
    for (int t = 0; t < n_frames; t++) {
        int ran = 0, was_reset = 0;
        const int id = cascade_step(c, models, pcm + (size_t)t * 160, results ? &results[t] : NULL, T, &ran, &was_reset);
        const nnsp_oracle_stream *s = &c->inst[id];
        if (valid) valid[t] = (int8_t)!was_reset;
        if (tap_logmel) memcpy(tap_logmel + (size_t)t * 40, s->feat.feature, 40 * sizeof(int32_t));
        if (tap_feat) { if (was_reset) memset(tap_feat + (size_t)t * 40, 0, 80); else memcpy(tap_feat + (size_t)t * 40, s->feat.ctx + 200, 80); }
        if (tap_h) { memset(tap_h + (size_t)t * 128, 0, 256); if (!was_reset) memcpy(tap_h + (size_t)t * 128, s->h, (size_t)model_h_stride(models[id]) * 2); }
        if (tap_c) { memset(tap_c + (size_t)t * 128, 0, 512); if (!was_reset) memcpy(tap_c + (size_t)t * 128, s->c, (size_t)model_h_stride(models[id]) * 4); }
        if (tap_post) {
            if (was_reset) { memset(tap_post + (size_t)t * 16, 0, 32); tap_post[(size_t)t * 16 + 15] = (int16_t)id; }
            else fill_post(tap_post + (size_t)t * 16, s, ran, id);
        }
    }
    return 0;
}

/* ======================================================================================== */
/* Multi-threaded batch runners (CPU baseline "port")                                         */
/* ======================================================================================== */
typedef struct {
    const nnsp_b200_model *m; const nnsp_b200_model *const *models; const int *seq; int len_seq;
    const nnsp_b200_cascade_params *params;
    int s0, s1, n_frames; const int16_t *pcm; long long stride; int16_t tp, tc;
    nnsp_b200_result *res; nnsp_b200_cascade_result *cres;
} job_t;

static void *job_nnsp(void *p)
{
    job_t *j = (job_t *)p;
    nnsp_oracle_stream *s = nnsp_oracle_stream_new();
    for (int i = j->s0; i < j->s1; i++)
        nnsp_oracle_nnsp_run(j->m, s, 1, j->pcm + (size_t)i * j->stride, j->n_frames, j->tp, j->tc,
                             j->res ? j->res + (size_t)i * j->n_frames : NULL, 0, 0, 0, 0, 0, 0, 0);
    nnsp_oracle_stream_free(s);
    return NULL;
}
static void *job_cascade(void *p)
{
    job_t *j = (job_t *)p;
    nnsp_oracle_cascade *c = nnsp_oracle_cascade_new();
    for (int i = j->s0; i < j->s1; i++)
        nnsp_oracle_cascade_run(j->models, c, 1, j->seq, j->len_seq, j->params, j->pcm + (size_t)i * j->stride,
                                j->n_frames, j->cres ? j->cres + (size_t)i * j->n_frames : NULL, 0, 0, 0, 0, 0, 0);
    nnsp_oracle_cascade_free(c);
    return NULL;
}

static double run_jobs(job_t *proto, int n_streams, int n_threads, void *(*fn)(void *))
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_streams) n_threads = n_streams;
    if (n_threads > 1024) n_threads = 1024;
    if (!nnsp_tables_get()) return -1.0;
    pthread_t th[1024];
    job_t jobs[1024];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < n_threads; k++) {
        jobs[k] = *proto;
        jobs[k].s0 = (int)((long long)n_streams * k / n_threads);
        jobs[k].s1 = (int)((long long)n_streams * (k + 1) / n_threads);
        if (pthread_create(&th[k], NULL, fn, &jobs[k]) != 0) return -1.0;
    }
    for (int k = 0; k < n_threads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double nnsp_oracle_batch_run(const nnsp_b200_model *m, int n_streams, const int16_t *pcm,
                             long long stream_stride, int n_frames, int16_t thresh_prob,
                             int16_t th_count, nnsp_b200_result *results, int n_threads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.m = m; j.pcm = pcm; j.stride = stream_stride; j.n_frames = n_frames; j.tp = thresh_prob; j.tc = th_count; j.res = results;
    return run_jobs(&j, n_streams, n_threads, job_nnsp);
}

double nnsp_oracle_cascade_batch_run(const nnsp_b200_model *const models[3], const int *seq,
                                     int len_seq, const nnsp_b200_cascade_params *params,
                                     int n_streams, const int16_t *pcm, long long stream_stride,
                                     int n_frames, nnsp_b200_cascade_result *results, int n_threads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.models = models; j.seq = seq; j.len_seq = len_seq; j.params = params; j.pcm = pcm; j.stride = stream_stride;
    j.n_frames = n_frames; j.cres = results;
    return run_jobs(&j, n_streams, n_threads, job_cascade);
}

/* ---- ingest (evb/src/main_nnsp.cc:58-65, audio_frame_callback) --------------------------------------- */
void nnsp_oracle_ingest_audadc(const uint32_t *raw, int16_t *pcm, long long n_frames)
{
    for (long long f = 0; f < n_frames; f++) {
        const uint32_t *r = raw + f * 160;
        int16_t *o = pcm + f * 160;
        for (int i = 0; i < 160; i++) {
            o[i] = (int16_t)(r[i] & 0x0000FFF0);                 /* :59 */
            if (i == 4) o[3] = (int16_t)((o[2] + o[4]) >> 1);       /* :61-64 */
        }
    }
}
