"""The C-ABI library loads without a GPU and exports every symbol its headers declare; struct layouts of the
legacy interface equal the reference's; with no device every compute entry point fails loudly."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from common import ROOT, have_reference_tree

LEGACY = ["NNSPClass_init", "NNSPClass_reset", "NNSPClass_exec", "FeatureClass_construct", "FeatureClass_setDefault",
          "FeatureClass_execute", "NeuralNetClass_init", "NeuralNetClass_setDefault", "NeuralNetClass_exe",
          "fc_8x16", "fc_8x16_acc32b", "lstm_8x16", "lstm_8x16_acc32b", "tanh_fix", "sigmoid_fix", "relu6_fix", "linear_fix"]


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnsp_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(nb):
    from nnsp_b200 import capi
    lib = capi.lib()
    declared = _declared("nnsp_b200.h")
    assert len(declared) >= 40
    for name in declared + LEGACY:
        assert hasattr(lib, name), "libnnsp_b200.so does not export %s" % name
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with include/nnsp_b200.h"


def test_result_record_layouts(nb):
    assert nb.RESULT_DT.itemsize == 8 and nb.CASCADE_RESULT_DT.itemsize == 12


def test_no_cpu_fallback(nb):
    if nb.device_count() > 0:
        pytest.skip("a GPU is present")
    m = nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad.nnspm"))
    with pytest.raises(nb.NnspError, match="no CPU fallback|CUDA"):
        nb.NNSPBatch(m, 8)
    with pytest.raises(nb.NnspError):
        nb.Cascade([nb.Model.from_blob(os.path.join(nb.MODEL_DIR, f)) for f in ("s2i.nnspm", "vad.nnspm", "kws_galaxy.nnspm")], 8)
    with pytest.raises(nb.NnspError):
        nb.feature_stages(np.zeros((1, 480), np.int16))
    with pytest.raises(nb.NnspError):
        nb.Group(m, 64, [0, 0])
    with pytest.raises(nb.NnspError):
        nb.net_eval(m, np.zeros((2, 240), np.int16))


def test_product_does_not_link_or_import_the_oracle():
    out = subprocess.run(["ldd", os.path.join(ROOT, "nnsp_b200", "libnnsp_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nnsp_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "nnsp_oracle" not in txt and "libnnsp_ref" not in txt, f


PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "neural_nets.h"
#include "feature_module.h"
#include "nn_speech.h"
int main(void) {
    printf("%zu %zu %zu %zu ", sizeof(NeuralNetClass), sizeof(FeatureClass), sizeof(NNSPClass), sizeof(stftModule));
    printf("%zu %zu %zu %zu %zu %zu ", offsetof(NeuralNetClass, size_layer), offsetof(NeuralNetClass, qbit_bias),
           offsetof(NeuralNetClass, pt_cstate), offsetof(NeuralNetClass, act_func), offsetof(NeuralNetClass, layer_func),
           offsetof(NeuralNetClass, pt_kernel_rec));
    printf("%zu %zu %zu %zu ", offsetof(FeatureClass, feature), offsetof(FeatureClass, normFeatContext),
           offsetof(FeatureClass, pt_norm_mean), offsetof(FeatureClass, qbit_output));
    printf("%zu %zu %zu %zu %zu\n", offsetof(NNSPClass, slides), offsetof(NNSPClass, pt_thresh_prob),
           offsetof(NNSPClass, counts_category), offsetof(NNSPClass, outputs), offsetof(NNSPClass, argmax_last));
    return 0;
}
"""


@pytest.mark.skipif(not have_reference_tree(), reason="reference headers not on this machine")
def test_legacy_struct_layouts_equal_the_reference_headers():
    outs = []
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe.c")
        open(src, "w").write(PROBE)
        for inc in ("/root/reference/ns-nnsp/includes-api", os.path.join(ROOT, "include", "nnsp_compat")):
            exe = os.path.join(d, "probe_" + str(len(outs)))
            subprocess.run(["gcc", "-I" + inc, src, "-o", exe], check=True)
            outs.append(subprocess.run([exe], capture_output=True, text=True).stdout)
    assert outs[0] == outs[1] and len(outs[0].split()) == 19
