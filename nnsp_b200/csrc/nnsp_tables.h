/* nnsp_tables.h -- constant tables of the ns-nnsp front end, generated (not copied).
 *
 * The reference ships these as C initialisers produced by its Python tool chain
 * (python/nnsp_pack/gen_stft_win.py, mel.py, fakefix_fft.py; LUTs in fixlog10.c:6 and
 * activation.c:5). nnsp-b200 recomputes them from their mathematical definitions with a
 * small libm-free double-precision kit (bit-reproducible on any IEEE-754 host) and checks
 * an FNV-1a fingerprint of every table before the engine is allowed to run.
 */
#ifndef NNSP_TABLES_H
#define NNSP_TABLES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NNSP_TBL_WIN_LEN     480
#define NNSP_TBL_FFT_TW_LEN  256   /* 64 butterflies x 4 packed complex16 */
#define NNSP_TBL_RFFT_TW_LEN 256
#define NNSP_TBL_BITREV_LEN  256
#define NNSP_TBL_MEL_LEN     534   /* 40 x {start,end} + 454 taps */
#define NNSP_TBL_MEL_TAPS    454
#define NNSP_TBL_LOG_LEN     256   /* 128 x {value, slope} */
#define NNSP_TBL_TANH_LEN    384   /* 192 x {value, slope} */

typedef struct {
    int16_t  stft_win[NNSP_TBL_WIN_LEN];       /* window_stft_coef.c:6   Q15 sqrt-Hann          */
    int32_t  fft_tw[NNSP_TBL_FFT_TW_LEN];      /* twiddle_fft_dif.c:8    lo16 = re, hi16 = im    */
    int32_t  rfft_tw[NNSP_TBL_RFFT_TW_LEN];    /* twiddle_fft_dif.c:74                           */
    int16_t  bitrev[NNSP_TBL_BITREV_LEN];      /* twiddle_fft_dif.c:76                           */
    int16_t  mel[NNSP_TBL_MEL_LEN];            /* melSpec_coeff.c:5                              */
    int16_t  log_lut[NNSP_TBL_LOG_LEN];        /* fixlog10.c:6                                   */
    int16_t  tanh_lut[NNSP_TBL_TANH_LEN];      /* activation.c:5                                 */
    /* derived, GPU-friendly views of the sparse mel table */
    int16_t  mel_start[40], mel_end[40], mel_off[40];  /* first/last bin, offset of first tap in mel_taps */
    int16_t  mel_taps[NNSP_TBL_MEL_TAPS];
} nnsp_tables;

/* Returns the process-wide table set (built once, thread-safe), or NULL if the
 * self-check fingerprint does not match (the engine then refuses to start). */
const nnsp_tables *nnsp_tables_get(void);
/* FNV-1a-64 of the seven primary tables, in declaration order. */
uint64_t nnsp_tables_fingerprint(const nnsp_tables *t);

#ifdef __cplusplus
}
#endif
#endif
