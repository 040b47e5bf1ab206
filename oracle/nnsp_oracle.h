/* oracle/nnsp_oracle.h -- TEST INFRASTRUCTURE. CPU restatement of the ns-nnsp hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; it is the checker, never the product (libnnsp_b200.so does not link
 * it and has no CPU path).
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py compares every function below with the
 * unmodified reference compiled by oracle/Makefile (oracle/_ref/) on the reference's own
 * test wavs and on synthetic/adversarial input, and tests/test_golden.py compares it with the
 * fixtures under tests/golden/ that were generated from that reference build.
 *
 * Unlike the reference (global scratch) it is re-entrant: all state
 * lives in nnsp_oracle_stream / nnsp_oracle_cascade objects, so streams can run on threads.
 */
#ifndef NNSP_ORACLE_H
#define NNSP_ORACLE_H
#include <stdint.h>
#include "nnsp_b200.h"      /* result / tap layouts, nnsp_b200_model (parsed by the product's blob reader) */
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nnsp_oracle_stream nnsp_oracle_stream;     /* one NNSPClass + FeatureClass + LSTM state */
typedef struct nnsp_oracle_cascade nnsp_oracle_cascade;   /* one nnCntrlClass + PcmBufClass + 3 instances */

nnsp_oracle_stream *nnsp_oracle_stream_new(void);
void nnsp_oracle_stream_free(nnsp_oracle_stream *s);

void nnsp_oracle_feature_stages(const int16_t *win480, int32_t *fft_in, int32_t *spec,
                                int32_t *pspec, int32_t *mel, int32_t *logmel);

/* taps: single-stream versions of nnsp_b200_taps rows, [T][...]; any may be NULL */
int nnsp_oracle_nnsp_run(const nnsp_b200_model *m, nnsp_oracle_stream *s, int do_reset,
                         const int16_t *pcm, int n_frames, int16_t thresh_prob, int16_t th_count,
                         nnsp_b200_result *results, int32_t *tap_logmel, int16_t *tap_feat,
                         int16_t *tap_act, int32_t *tap_logits, int16_t *tap_h, int32_t *tap_c,
                         int16_t *tap_post);

/* one network evaluation on an explicit input vector and explicit LSTM state (unit tests of
 * the layer arithmetic: saturation, wrap, Q-format corner cases) */
int nnsp_oracle_net_eval(const nnsp_b200_model *m, const int16_t *input, int16_t *h_inout,
                         int32_t *c_inout, int16_t *act, int32_t *logits);

/* evb/src/main_nnsp.cc:58-65: raw AUDADC words -> conditioned int16 PCM, frame by frame */
void nnsp_oracle_ingest_audadc(const uint32_t *raw, int16_t *pcm, long long n_frames);

void nnsp_oracle_default_params(nnsp_b200_cascade_params *p);
nnsp_oracle_cascade *nnsp_oracle_cascade_new(void);
void nnsp_oracle_cascade_free(nnsp_oracle_cascade *c);
int nnsp_oracle_cascade_run(const nnsp_b200_model *const models[3], nnsp_oracle_cascade *c,
                            int do_reset, const int *seq, int len_seq,
                            const nnsp_b200_cascade_params *params, const int16_t *pcm,
                            int n_frames, nnsp_b200_cascade_result *results, int32_t *tap_logmel,
                            int16_t *tap_feat, int16_t *tap_h, int32_t *tap_c, int16_t *tap_post,
                            int8_t *valid);

/* Many independent streams on n_threads host threads (the CPU baseline "port"):
 * stream s reads pcm + s*stream_stride; results [n_streams][n_frames] (may be NULL).
 * Returns elapsed seconds of the frame loops (CLOCK_MONOTONIC), < 0 on error. */
double nnsp_oracle_batch_run(const nnsp_b200_model *m, int n_streams, const int16_t *pcm,
                             long long stream_stride, int n_frames, int16_t thresh_prob,
                             int16_t th_count, nnsp_b200_result *results, int n_threads);
double nnsp_oracle_cascade_batch_run(const nnsp_b200_model *const models[3], const int *seq,
                                     int len_seq, const nnsp_b200_cascade_params *params,
                                     int n_streams, const int16_t *pcm, long long stream_stride,
                                     int n_frames, nnsp_b200_cascade_result *results, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
