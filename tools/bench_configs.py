#!/usr/bin/env python
"""Device-resident timing of other BASELINE.json configurations (not the bench.py headline):
    python tools/bench_configs.py [vad:4096 kws:16384:acc32 s2i:32768 cascade:8192 ...] [--paths imma,dp2a]
Prints one JSON line per (config, network kernel)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import nnsp_b200 as nb  # noqa: E402

FILES = {"s2i": "s2i.nnspm", "vad": "vad.nnspm", "kws": "kws_galaxy.nnspm"}


def time_handle(h, exec_fn, T, S, steps=8, warm=3):
    ev0, ev1 = nb.Event(), nb.Event()
    for i in range(warm):
        exec_fn(i)
    h.sync()
    ev0.record(h.stream)
    for i in range(steps):
        exec_fn(i)
    h.sync()                       # every stream of the handle (calls are pipelined over two)
    ev1.record(h.stream)
    h.sync()
    ms = ev0.elapsed_ms_to(ev1) / steps
    exec_fn(0)
    h.sync()
    return ms, h.last_kernel_ms()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    paths = ["split", "imma"]
    cpaths = ["sorted", "sequential"]
    for a in sys.argv[1:]:
        if a.startswith("--paths"):
            paths = a.split("=")[1].split(",")
        if a.startswith("--cascade-paths"):
            cpaths = a.split("=")[1].split(",")
    if not args:
        args = ["vad:4096", "kws:16384:acc32", "s2i:32768", "cascade:8192"]
    T = 100
    for spec in args:
        parts = spec.split(":")
        name, S = parts[0], int(parts[1])
        acc32 = "acc32" in parts[2:]
        pool = min(S, 2048)
        base = nb.synth_pcm(pool, T)
        pcm = np.tile(base, ((S + pool - 1) // pool, 1))[:S]
        d = [nb.DeviceArray.from_host(pcm), nb.DeviceArray.from_host(np.roll(pcm, 7, axis=0))]
        if name == "cascade":
            models = [nb.Model.from_blob(os.path.join(nb.MODEL_DIR, FILES[k]), acc32=acc32) for k in ("s2i", "vad", "kws")]
            for cp in cpaths:
                h = nb.Cascade(models, S)
                h.set_path(cp)
                res = nb.DeviceArray((S, T), nb.CASCADE_RESULT_DT)
                ms, km = time_handle(h, lambda i: h.exec_device(d[i & 1], T * 160, T, res), T, S, steps=20, warm=10)
                stage = np.bincount(res.to_host()["stage_id"].ravel().astype(np.int64), minlength=3)
                print(json.dumps({"config": spec, "lib": os.path.basename(os.environ.get("NNSP_B200_LIB", "product")), "kernel": "cascade-" + cp, "ms_per_step": ms, "audio_s_per_s": S * T * 0.01 / (ms * 1e-3),
                                  "feat_ms": km[0], "nn_ms": km[1], "stage_frames_last_step": stage.tolist()}), flush=True)
                h.close()
                res.free()
        else:
            m = nb.Model.from_blob(os.path.join(nb.MODEL_DIR, FILES[name]), acc32=acc32)
            for p in paths:
                h = nb.NNSPBatch(m, S, nn_path=p)
                res = nb.DeviceArray((S, T), nb.RESULT_DT)
                ms, km = time_handle(h, lambda i: h.exec_device(d[i & 1], T * 160, T, res), T, S)
                print(json.dumps({"config": spec, "lib": os.path.basename(os.environ.get("NNSP_B200_LIB", "product")), "kernel": p, "ms_per_step": ms, "audio_s_per_s": S * T * 0.01 / (ms * 1e-3),
                                  "feat_ms": km[0], "nn_ms": km[1]}), flush=True)
                h.close()
                res.free()
        for x in d:
            x.free()


if __name__ == "__main__":
    main()
