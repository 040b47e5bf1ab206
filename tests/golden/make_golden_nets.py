#!/usr/bin/env python
"""Generate tests/golden/golden_nets_v1.npz: NeuralNetClass_exe of the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) on SYNTHETIC and CRAFTED layer stacks with adversarial inputs and LSTM states.

The model tables come from tests/common.py (NET_CASES, NET_CRAFTED -> make_blob, deterministic); the reference reads
them through a NeuralNetClass literal the glue builds over the container (oracle/ref_glue.c ref_net_eval_blob), once
with fc_8x16 / lstm_8x16 and once with the _acc32b twins. What the crafted cases reach (reference lines):
  shift_32b saturation of the lstm input half ........ affine_acc32b.c:349-407 (rc_Krows), :566-592
  bias << 30 wrapping modulo 2^32 ..................... affine_acc32b.c:190-217
  tanh_fix / sigmoid_fix beyond |x| >= 5.0 ............ activation.c:31-86
  cell state saturating to int32 ...................... lstm.c:106-115
Run only where /root/reference exists; the fixture itself travels. Usage: python tests/golden/make_golden_nets.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import NET_CASES, NET_CRAFTED, adversarial_net_inputs, adversarial_states, net_case_blob, net_case_dims  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402


def main():
    R = RefLib(False)                     # both accumulator flavours of the network units are linked into either library
    out = {}
    rng = np.random.default_rng(23)
    xs = adversarial_net_inputs(rng)
    out["x"] = xs
    for case in NET_CASES + NET_CRAFTED:
        name = case[0]
        a_s, h_s, n_o = net_case_dims(case)
        for acc32 in (False, True):
            blob = net_case_blob(case, acc32)
            h0, c0 = adversarial_states(rng, len(xs), h_s)
            acts, logs, h1, c1 = [], [], [], []
            for i in range(len(xs)):
                a, l, hh, cc = R.net_eval_blob(blob, acc32, xs[i], h0[i, :h_s], c0[i, :h_s], a_s, n_o)
                acts.append(a.copy()); logs.append(l.copy()); h1.append(hh.copy()); c1.append(cc.copy())
            key = "%s_%s" % (name, "acc32" if acc32 else "acc64")
            out[key + "_blob_sha"] = np.array([hashlib.sha256(blob).hexdigest()])
            out[key + "_h0"] = h0; out[key + "_c0"] = c0
            out[key + "_act"] = np.stack(acts); out[key + "_logits"] = np.stack(logs)
            out[key + "_h1"] = np.stack(h1) if h_s else np.zeros((len(xs), 0), np.int16)
            out[key + "_c1"] = np.stack(c1) if h_s else np.zeros((len(xs), 0), np.int32)
    path = os.path.join(HERE, "golden_nets_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
