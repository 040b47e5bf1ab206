/* nnsp_engine.cuh -- device-visible data structures of the nnsp-b200 engine. */
#pragma once
#include <stdint.h>
#include "nnsp_b200.h"

namespace nnsp {

/* ---- constant tables as the kernels see them (global memory, copied to SMEM per CTA) ---- */
constexpr int MEL_GROUPS = 144;
/* most groups of any band in mel round r (bands 39-16r .. 24-16r, widest first): bands 24..39, 8..23, 0..7 */
constexpr int MEL_MAXG0 = 8, MEL_MAXG1 = 4, MEL_MAXG2 = 2;    /* 4-bin groups of the 40 mel bands, each band padded to aligned groups */

struct DevTables {
    /* mel filterbank (melSpec_coeff.c:5) regrouped for 128-bit loads: band b owns groups g0 .. g0+ng-1, group i
     * holds the taps of bins bin0+4i .. bin0+4i+3 (bin0 = first bin rounded down to 4; taps outside the band are 0);
     * mel_meta[b] = g0 | ng << 8 | bin0 << 16 */
    int4     mel_tap4[MEL_GROUPS];
    int2     win2[240];        /* stft window, sign-extended Q15 coefficients of samples 2p, 2p+1 */
    /* radix-4 twiddles, sign-extended (re, im), laid out in the order the half-warp reads them:
     *   tw0[a][n][L] = tw^(n+1 column) of butterfly k = L + 16a (stage 0), tw1[n][L] of k = 4L (stage 1),
     *   tw2[m][n] of k = 16m (stage 2, same for every lane); column order tw^2, tw^1, tw^3 (fft.c:182) */
    int2     tw0[4][3][16];
    int2     tw1[3][16];
    int2     tw2[4][3];
    int2     rtw[257];         /* real-FFT split twiddles exp(-2 pi j k / 512), k = 0..256 (k = 256 unused) */
    uint32_t mel_meta[40];
    int16_t  log_lut[256];
    int16_t  tanh_lut[384];
    /* tanh_fix (activation.c:31-69) folded for the vector kernels: for w = min(|x| >> 9, 320), d = |x| - 512 w,
     * tanh_fix(|x|) = (d * tanh2[w].x + tanh2[w].y) >> 15 -- value, the clamp at 0 and the saturation at 5.0 are in the
     * constants; fill_dev_tables checks every |x| up to the saturation against the definition */
    int2     tanh2[321];
};

/* ---- one layer, GPU layout -------------------------------------------------------------
 * Weights are stored K-major in 32-bit words: word (k4, r) holds W[r][4*k4 .. 4*k4+3]
 * (int8, little endian), rows padded to a multiple of 32 with zero rows, K padded to a
 * multiple of 4 with zero columns. lstm rows are ordered gate*H + unit (i, j, f, o). */
struct DevLayer {
    int type, act;
    int rows;            /* units out (H for lstm)                 */
    int nrows;           /* rows of W: rows (fc) or 4*rows (lstm)  */
    int nrows_pad;       /* nrows rounded up to 32                 */
    int cols, k4;        /* inputs, ceil(cols/4)                   */
    int k4rec;           /* lstm: ceil(rows/4)                     */
    int acc32;
    int sh_x;            /* lstm: qi_next - qi          (affine.c:371)         */
    int sh_bias;         /* qs - qbit_bias              (affine.c:192)         */
    int sh_out;          /* 15 - qs                     (affine.c:244)         */
    int w_off, wrec_off; /* word offsets into the weight image                 */
    int bias_off;        /* int16 offset into the bias image                   */
};

struct DevModel {
    int      nn_id, numlayers;
    int      act_stride, h_stride, n_out;
    int      feat_rshift;                   /* 30 - qbit_input[0] (feature_module.c:70) */
    int      weight_words, bias_count;      /* sizes of the two images                  */
    int32_t  mean[40], stdR[40];
    int16_t  silence[40];                   /* standardised log10(2^-15) row (feature_module.c:32-43) */
    DevLayer layer[NNSP_B200_MAX_LAYERS];
};

/* per-stream NNSPClass scalars, same order as the `post` tap */
enum { SC_TRIGGER = 0, SC_OUT0 = 1, SC_CNT0 = 4, SC_ARGMAX_LAST = 12, SC_SLIDES = 13, SC_RAN = 14, SC_STAGE = 15, SC_N = 16 };

/* cascade-only per-stream scalars */
enum { CS_POS = 0, CS_CNT_KWS = 1, CS_CNT_S2I = 2, CS_AGE = 3, CS_N = 4 };

struct StreamState {           /* all arrays stream-major, dense */
    int16_t *ctx;              /* [S][240]  normFeatContext, 6 rows x 40 (row 5 survives a reset) */
    int16_t *h;                /* [S][h_stride]                                        */
    int32_t *c;                /* [S][h_stride]                                        */
    int16_t *scal;             /* [S][16]                                              */
    int16_t *hist;             /* [S][hist_frames*160] newest PCM frames of the previous exec call */
    int32_t *lmhist;           /* cascade: [S][lm_hist][40] newest log-mel rows of previous calls  */
    uint16_t *casc;            /* cascade: [S][4]                                      */
};

}  // namespace nnsp
