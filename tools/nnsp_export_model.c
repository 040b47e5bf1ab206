/* tools/nnsp_export_model.c -- turn a generated model table (def_nn<id>_<name>.c, written by
 * the reference's python/c_code_table_converter.py) into an NNSPM1 blob that
 * nnsp_b200_model_from_blob loads at run time.
 *
 * The table file is compiled UNMODIFIED against include/nnsp_compat (see tools/Makefile):
 *   gcc -DNET=net_vad -DMEAN=feature_mean_vad -DSTDR=feature_stdR_vad -DNN_ID=1 \
 *       -Iinclude -Iinclude/nnsp_compat tools/nnsp_export_model.c path/to/def_nn1_vad.c \
 *       nnsp_b200/csrc/nnsp_model.c nnsp_b200/csrc/nnsp_model_net.c nnsp_b200/csrc/nnsp_legacy_tags.c
 * Add -DDEF_ACC32BIT_OPT to export the wrapping-32-bit-accumulator flavour. */
#include <stdio.h>
#include <stdlib.h>
#include "nnsp_b200.h"
#include "nnsp_compat/nnsp_legacy_api.h"

extern NeuralNetClass NET;
extern const int32_t MEAN[];
extern const int32_t STDR[];

int main(int argc, char **argv)
{
    if (argc != 2) { fprintf(stderr, "usage: %s out.nnspm\n", argv[0]); return 2; }
    nnsp_b200_model *m = NULL;
    int rc = nnsp_b200_model_from_net(&NET, MEAN, STDR, NN_ID, &m);
    if (rc) { fprintf(stderr, "model_from_net: %s (%s)\n", nnsp_b200_strerror(rc), nnsp_b200_last_error()); return 1; }
    size_t n = 0;
    nnsp_b200_model_to_blob(m, NULL, 0, &n);
    void *buf = malloc(n);
    rc = nnsp_b200_model_to_blob(m, buf, n, &n);
    if (rc) { fprintf(stderr, "model_to_blob: %s\n", nnsp_b200_strerror(rc)); return 1; }
    FILE *f = fopen(argv[1], "wb");
    if (!f || fwrite(buf, 1, n, f) != n) { perror(argv[1]); return 1; }
    fclose(f);
    int acc32 = 0, nl = 0;
    nnsp_b200_model_info(m, NULL, &nl, NULL, &acc32);
    printf("%s: %zu bytes, %d layers, acc32=%d\n", argv[1], n, nl, acc32);
    nnsp_b200_model_free(m);
    free(buf);
    return 0;
}
