/* nnsp_tables.c -- see nnsp_tables.h. Compile with -ffp-contract=off (no FMA fusion) so the
 * double arithmetic below is bit-reproducible on every IEEE-754 host. Only + - * / and
 * sqrt (correctly rounded by IEEE) are used; no libm transcendental is called. */
#include "nnsp_tables.h"
#include <math.h>     /* sqrt, floor only */
#include <pthread.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* libm-free math kit                                                                    */
/* ------------------------------------------------------------------------------------ */
static const double DM_PI   = 3.14159265358979323846264338327950288;
static const double DM_LN2  = 0.693147180559945309417232121458176568;
static const double DM_LN10 = 2.30258509299404568401799145468436421;

/* sin and cos of r, 0 <= r <= pi/4 (Taylor, terms below 1e-19) */
static void dm_sincos_small(double r, double *s, double *c)
{
    double r2 = r * r, ts = r, tc = 1.0, ss = r, cs = 1.0;
    for (int n = 1; n <= 12; n++) {
        tc = -tc * r2 / (double)((2 * n - 1) * (2 * n));
        ts = -ts * r2 / (double)((2 * n) * (2 * n + 1));
        cs += tc;
        ss += ts;
    }
    *s = ss;
    *c = cs;
}

/* sin/cos of 2*pi*num/den with exact octant symmetries (so cos(pi/2) is exactly 0) */
static void dm_sincos_turn(long num, long den, double *s, double *c)
{
    num %= den;
    if (num < 0) num += den;
    /* position inside the turn in eighths: oct = floor(8*num/den), rem/den8 in [0,1) of an octant */
    long n8 = num * 8;
    long oct = n8 / den, rem = n8 % den;       /* angle = (oct + rem/den) * pi/4 */
    double a, sa, ca;
    if (oct & 1) {                              /* mirror odd octants onto [0, pi/4] */
        a = (double)(den - rem) / (double)den * (DM_PI / 4.0);
    } else {
        a = (double)rem / (double)den * (DM_PI / 4.0);
    }
    if (rem == 0 && (oct & 1)) { sa = 0.70710678118654752440; ca = sa; } /* exact 45 deg keeps symmetry */
    else dm_sincos_small(a, &sa, &ca);
    if (rem == 0 && !(oct & 1)) { sa = 0.0; ca = 1.0; }
    double ss, cc;
    switch (oct) {
    case 0: ss = sa;  cc = ca;  break;
    case 1: ss = ca;  cc = sa;  break;
    case 2: ss = ca;  cc = -sa; break;
    case 3: ss = sa;  cc = -ca; break;
    case 4: ss = -sa; cc = -ca; break;
    case 5: ss = -ca; cc = -sa; break;
    case 6: ss = -ca; cc = sa;  break;
    default: ss = -sa; cc = ca; break;
    }
    *s = ss;
    *c = cc;
}

/* natural log, x > 0 */
static double dm_log(double x)
{
    int e = 0;
    while (x >= 1.4142135623730951) { x *= 0.5; e++; }
    while (x < 0.70710678118654757) { x *= 2.0; e--; }
    double z = (x - 1.0) / (x + 1.0), z2 = z * z, t = z, sum = 0.0;
    for (int n = 0; n < 30; n++) {
        sum += t / (double)(2 * n + 1);
        t *= z2;
    }
    return 2.0 * sum + (double)e * DM_LN2;
}

static double dm_exp(double x)
{
    double kf = floor(x / DM_LN2 + 0.5);
    int k = (int)kf;
    double r = x - kf * DM_LN2, t = 1.0, sum = 1.0;
    for (int n = 1; n <= 26; n++) {
        t = t * r / (double)n;
        sum += t;
    }
    while (k > 0) { sum *= 2.0; k--; }
    while (k < 0) { sum *= 0.5; k++; }
    return sum;
}

static double dm_tanh(double x)   /* x >= 0 */
{
    double e2 = dm_exp(2.0 * x);
    return (e2 - 1.0) / (e2 + 1.0);
}

/* python/nnsp_pack/converter_fix_point.py:7-15 `fakefix(v,16,15)` followed by *2^15: floor + clamp */
static int32_t q15_floor(double v)
{
    double f = floor(v * 32768.0);
    if (f > 32767.0) f = 32767.0;
    if (f < -32768.0) f = -32768.0;
    return (int32_t)f;
}

/* ------------------------------------------------------------------------------------ */
/* table generators                                                                      */
/* ------------------------------------------------------------------------------------ */
/* python/nnsp_pack/gen_stft_win.py:9-19: sqrt(hop/win * (1 - cos(2 pi i / win))) */
static void gen_window(int16_t *w)
{
    for (int i = 0; i < NNSP_TBL_WIN_LEN; i++) {
        double s, c;
        dm_sincos_turn(i, NNSP_TBL_WIN_LEN, &s, &c);
        double sq = (160.0 / 480.0) * (1.0 - c);
        w[i] = (int16_t)q15_floor(sqrt(sq));
    }
}

static int32_t pack_c16(double re, double im)
{
    uint32_t r = (uint16_t)(int16_t)q15_floor(re), i = (uint16_t)(int16_t)q15_floor(im);
    return (int32_t)(r | (i << 16));
}

/* python/nnsp_pack/fakefix_fft.py:34-43,73-82: tw = exp(-2j pi k / N); columns tw^0, tw^2, tw^1, tw^3 */
static void gen_twiddles(int32_t *fft_tw, int32_t *rfft_tw, int16_t *bitrev)
{
    static const int pw[4] = { 0, 2, 1, 3 };
    for (int k = 0; k < 64; k++)
        for (int j = 0; j < 4; j++) {
            double s, c;
            dm_sincos_turn((long)k * pw[j], 256, &s, &c);
            fft_tw[4 * k + j] = pack_c16(c, -s);
        }
    for (int k = 0; k < 256; k++) {
        double s, c;
        dm_sincos_turn(k, 512, &s, &c);
        rfft_tw[k] = pack_c16(c, -s);
    }
    for (int i = 0; i < 256; i++) {
        int r = 0;
        for (int b = 0; b < 8; b++)
            if (i & (1 << b)) r |= 1 << (7 - b);
        bitrev[i] = (int16_t)r;
    }
}

/* python/nnsp_pack/mel.py:11-52 (HTK mel scale, 40 triangles over a 512-point FFT at 16 kHz) */
static int gen_mel(nnsp_tables *t)
{
    enum { NF = 40 };
    double binm[NF + 2];
    double high = 2595.0 * (dm_log(1.0 + 8000.0 / 700.0) / DM_LN10);
    double step = (high - 0.0) / (double)(NF + 1);
    for (int i = 0; i < NF + 2; i++) {
        double mel = (i == NF + 1) ? high : 0.0 + (double)i * step;   /* numpy.linspace */
        double hz = 700.0 * (dm_exp(mel / 2595.0 * DM_LN10) - 1.0);
        binm[i] = floor((512.0 + 1.0) * hz / 16000.0);
    }
    int pos = 0, tap = 0;
    for (int m = 1; m <= NF; m++) {
        int lo = (int)binm[m - 1], ce = (int)binm[m], hi = (int)binm[m + 1];
        double row[257];
        memset(row, 0, sizeof row);
        for (int k = lo; k < ce; k++) row[k] = ((double)k - binm[m - 1]) / (binm[m] - binm[m - 1]);
        for (int k = ce; k < hi; k++) row[k] = (binm[m + 1] - (double)k) / (binm[m + 1] - binm[m]);
        if (pos + 2 + (hi - 1 - lo) > NNSP_TBL_MEL_LEN) return -1;
        t->mel[pos++] = (int16_t)(lo + 1);
        t->mel[pos++] = (int16_t)(hi - 1);
        t->mel_start[m - 1] = (int16_t)(lo + 1);
        t->mel_end[m - 1] = (int16_t)(hi - 1);
        t->mel_off[m - 1] = (int16_t)tap;
        for (int k = lo + 1; k < hi; k++) {
            int16_t v = (int16_t)q15_floor(row[k]);
            if (tap >= NNSP_TBL_MEL_TAPS) return -1;
            t->mel[pos++] = v;
            t->mel_taps[tap++] = v;
        }
    }
    return (pos == NNSP_TBL_MEL_LEN && tap == NNSP_TBL_MEL_TAPS) ? 0 : -1;
}

/* fixlog10.c:43-45: segment k of [1,2): value ln(1+k/128), slope 1/(1+k/128), Q15 floor */
static void gen_log_lut(int16_t *lut)
{
    for (int k = 0; k < 128; k++) {
        double x0 = 1.0 + (double)k / 128.0;
        lut[2 * k] = (int16_t)q15_floor(dm_log(x0));
        lut[2 * k + 1] = (int16_t)q15_floor(1.0 / x0);
    }
}

/* activation.c:58-60: segment k expands tanh around x0 = 1/64 + k/32: value, 1 - tanh^2 */
static void gen_tanh_lut(int16_t *lut)
{
    for (int k = 0; k < 192; k++) {
        double x0 = 1.0 / 64.0 + (double)k / 32.0, th = dm_tanh(x0);
        lut[2 * k] = (int16_t)q15_floor(th);
        lut[2 * k + 1] = (int16_t)q15_floor(1.0 - th * th);
    }
}

/* ------------------------------------------------------------------------------------ */
static uint64_t fnv1a(uint64_t h, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

uint64_t nnsp_tables_fingerprint(const nnsp_tables *t)
{
    uint64_t h = 0xcbf29ce484222325ULL;
    h = fnv1a(h, t->stft_win, sizeof t->stft_win);
    h = fnv1a(h, t->fft_tw, sizeof t->fft_tw);
    h = fnv1a(h, t->rfft_tw, sizeof t->rfft_tw);
    h = fnv1a(h, t->bitrev, sizeof t->bitrev);
    h = fnv1a(h, t->mel, sizeof t->mel);
    h = fnv1a(h, t->log_lut, sizeof t->log_lut);
    h = fnv1a(h, t->tanh_lut, sizeof t->tanh_lut);
    return h;
}

/* Fingerprint of the tables when they equal the reference's (verified against the reference
 * objects by tests/test_tables_and_format.py). */
#define NNSP_TABLES_EXPECTED_FNV 0x64c7c19ae737264dULL

static nnsp_tables g_tables;
static int g_ok = 0;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build_tables(void)
{
    memset(&g_tables, 0, sizeof g_tables);
    gen_window(g_tables.stft_win);
    gen_twiddles(g_tables.fft_tw, g_tables.rfft_tw, g_tables.bitrev);
    if (gen_mel(&g_tables) != 0) return;
    gen_log_lut(g_tables.log_lut);
    gen_tanh_lut(g_tables.tanh_lut);
    g_ok =
           (nnsp_tables_fingerprint(&g_tables) == NNSP_TABLES_EXPECTED_FNV);
}

const nnsp_tables *nnsp_tables_get(void)
{
    pthread_once(&g_once, build_tables);
    return g_ok ? &g_tables : NULL;
}
