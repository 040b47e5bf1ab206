"""oracle/pyoracle.py -- TEST INFRASTRUCTURE: ctypes access to the two CPU checkers.

* ``Oracle``  : oracle/libnnsp_oracle.so, the re-entrant C restatement (oracle/nnsp_oracle.c).
* ``RefLib``  : oracle/_ref/libnnsp_ref_acc{64,32}.so, the UNMODIFIED reference compiled by
                oracle/Makefile (single instance, global state: one stream at a time).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. The product (nnsp_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MODEL_DIR = os.path.join(ROOT, "nnsp_b200", "models")

S2I, VAD, KWS = 0, 1, 2
MODEL_FILES = {S2I: "s2i.nnspm", VAD: "vad.nnspm", KWS: "kws_galaxy.nnspm"}

RESULT_DT = np.dtype([("trigger", "<i2"), ("outputs", "<i2", (3,))])
CASCADE_RESULT_DT = np.dtype([("stage_id", "i1"), ("pos_after", "i1"), ("detected", "<i2"),
                              ("outputs", "<i2", (3,)), ("cnt_timeout", "<u2")])
assert RESULT_DT.itemsize == 8 and CASCADE_RESULT_DT.itemsize == 12


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build(ref=True):
    """(Re)build the checkers. The _ref build needs /root/reference and is skipped without it."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class Taps:
    """Host arrays in the tap layout of include/nnsp_b200.h for one stream and T frames."""

    def __init__(self, T, act_stride, h_stride, n_out):
        self.logmel = np.zeros((T, 40), np.int32)
        self.feat = np.zeros((T, 40), np.int16)
        self.act = np.zeros((T, act_stride), np.int16)
        self.logits = np.zeros((T, n_out), np.int32)
        self.h = np.zeros((T, h_stride), np.int16)
        self.c = np.zeros((T, h_stride), np.int32)
        self.post = np.zeros((T, 16), np.int16)

    def names(self):
        return ["logmel", "feat", "act", "logits", "h", "c", "post"]


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libnnsp_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.nnsp_b200_model_from_blob.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.nnsp_b200_model_set_acc32.argtypes = [C.c_void_p, C.c_int]
        L.nnsp_b200_model_free.argtypes = [C.c_void_p]
        L.nnsp_b200_model_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.c_void_p, C.POINTER(C.c_int)]
        L.nnsp_oracle_stream_new.restype = C.c_void_p
        L.nnsp_oracle_stream_free.argtypes = [C.c_void_p]
        L.nnsp_oracle_cascade_new.restype = C.c_void_p
        L.nnsp_oracle_cascade_free.argtypes = [C.c_void_p]
        L.nnsp_oracle_feature_stages.argtypes = [C.c_void_p] * 6
        L.nnsp_oracle_nnsp_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int16, C.c_int16] + [C.c_void_p] * 8
        L.nnsp_oracle_net_eval.argtypes = [C.c_void_p] * 6
        L.nnsp_oracle_cascade_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.nnsp_oracle_batch_run.restype = C.c_double
        L.nnsp_oracle_batch_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int,
                                            C.c_int16, C.c_int16, C.c_void_p, C.c_int]
        L.nnsp_oracle_cascade_batch_run.restype = C.c_double
        L.nnsp_oracle_cascade_batch_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                    C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int]
        L.nnsp_oracle_default_params.argtypes = [C.c_void_p]
        L.nnsp_oracle_ingest_audadc.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        self._models = {}

    # -- models ---------------------------------------------------------------------------
    def load_model(self, blob_bytes, acc32=None):
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(blob_bytes), len(blob_bytes))
        rc = self.lib.nnsp_b200_model_from_blob(buf, len(blob_bytes), C.byref(h))
        if rc:
            raise RuntimeError("model_from_blob rc=%d" % rc)
        if acc32 is not None:
            self.lib.nnsp_b200_model_set_acc32(h, int(acc32))
        return h

    def model(self, nn_id, acc32=False):
        key = (nn_id, bool(acc32))
        if key not in self._models:
            with open(os.path.join(MODEL_DIR, MODEL_FILES[nn_id]), "rb") as f:
                self._models[key] = self.load_model(f.read(), acc32)
        return self._models[key]

    def model_dims(self, m):
        nl = C.c_int()
        sizes = (C.c_int16 * 11)()
        self.lib.nnsp_b200_model_info(m, None, C.byref(nl), sizes, None)
        sz = list(sizes)[: nl.value + 1]
        # every shipped model has exactly one lstm (layer 1); h_stride is reported by the run itself
        return sum(sz[1:-1]), sz[-1], sz

    # -- front end --------------------------------------------------------------------------
    def feature_stages(self, win480):
        win480 = np.ascontiguousarray(win480, np.int16)
        assert win480.shape == (480,)
        out = dict(fft_in=np.zeros(512, np.int32), spec=np.zeros(514, np.int32),
                   pspec=np.zeros(257, np.int32), mel=np.zeros(40, np.int32), logmel=np.zeros(40, np.int32))
        self.lib.nnsp_oracle_feature_stages(_p(win480), _p(out["fft_in"]), _p(out["spec"]),
                                            _p(out["pspec"]), _p(out["mel"]), _p(out["logmel"]))
        return out

    def ingest_audadc(self, raw):
        raw = np.ascontiguousarray(raw, np.uint32)
        out = np.zeros(raw.shape, np.int16)
        self.lib.nnsp_oracle_ingest_audadc(_p(raw), _p(out), raw.size // 160)
        return out

    # -- one stream ---------------------------------------------------------------------------
    def nnsp_run(self, m, pcm, thresh_prob=16383, th_count=4, state=None, reset=True, h_stride=None, taps=True):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = len(pcm) // 160
        act_stride, n_out, sz = self.model_dims(m)
        if h_stride is None:
            h_stride = sz[2]
        own = state is None
        st = state or self.lib.nnsp_oracle_stream_new()
        res = np.zeros(T, RESULT_DT)
        tp = Taps(T, act_stride, h_stride, n_out) if taps else None
        a = [_p(getattr(tp, n)) for n in tp.names()] if taps else [None] * 7
        rc = self.lib.nnsp_oracle_nnsp_run(m, st, int(reset), _p(pcm), T, thresh_prob, th_count, _p(res), *a)   # reset: 0 continue, 1 new instance, 2 NNSPClass_reset
        if own:
            self.lib.nnsp_oracle_stream_free(st)
        if rc:
            raise RuntimeError("oracle nnsp_run rc=%d" % rc)
        return res, tp

    def net_eval(self, m, x, h, c):
        act_stride, n_out, sz = self.model_dims(m)
        x = np.ascontiguousarray(x, np.int16)
        h = np.ascontiguousarray(h, np.int16).copy()
        c = np.ascontiguousarray(c, np.int32).copy()
        act = np.zeros(act_stride, np.int16)
        logits = np.zeros(n_out, np.int32)
        self.lib.nnsp_oracle_net_eval(m, _p(x), _p(h), _p(c), _p(act), _p(logits))
        return act, logits, h, c

    # -- cascade ------------------------------------------------------------------------------
    def default_params(self):
        p = np.zeros(10, np.int16)
        self.lib.nnsp_oracle_default_params(_p(p))
        return p

    def _model_array(self, models):
        arr = (C.c_void_p * 3)()
        for i in range(3):
            arr[i] = models[i]
        return arr

    def cascade_run(self, models, pcm, seq=(VAD, KWS, S2I), params=None, state=None, reset=True, taps=True):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = len(pcm) // 160
        own = state is None
        st = state or self.lib.nnsp_oracle_cascade_new()
        seq_a = np.asarray(seq, np.int32)
        # reset=False with params: the application rewrites the live controller's Params (seen by the next frame); without: kept
        par = (None if (not reset or reset == 2) else self.default_params()) if params is None else np.ascontiguousarray(params, np.int16)
        res = np.zeros(T, CASCADE_RESULT_DT)
        tp = Taps(T, 1, 128, 1) if taps else None
        valid = np.zeros(T, np.int8)
        a = [_p(tp.logmel), _p(tp.feat), _p(tp.h), _p(tp.c), _p(tp.post)] if taps else [None] * 5
        rc = self.lib.nnsp_oracle_cascade_run(self._model_array(models), st, int(reset), _p(seq_a), len(seq_a),
                                              _p(par), _p(pcm), T, _p(res), *a, _p(valid))
        if own:
            self.lib.nnsp_oracle_cascade_free(st)
        if rc:
            raise RuntimeError("oracle cascade_run rc=%d" % rc)
        return res, tp, valid

    # -- throughput (CPU baseline "port") -----------------------------------------------------
    def batch_run(self, m, pcm2d, thresh_prob=16383, th_count=4, n_threads=1, want_results=False):
        pcm2d = np.ascontiguousarray(pcm2d, np.int16)
        S, n = pcm2d.shape
        T = n // 160
        res = np.zeros((S, T), RESULT_DT) if want_results else None
        sec = self.lib.nnsp_oracle_batch_run(m, S, _p(pcm2d), n, T, thresh_prob, th_count, _p(res), n_threads)
        return sec, res

    def cascade_batch_run(self, models, pcm2d, seq=(VAD, KWS, S2I), params=None, n_threads=1, want_results=False):
        pcm2d = np.ascontiguousarray(pcm2d, np.int16)
        S, n = pcm2d.shape
        T = n // 160
        seq_a = np.asarray(seq, np.int32)
        par = self.default_params() if params is None else np.ascontiguousarray(params, np.int16)
        res = np.zeros((S, T), CASCADE_RESULT_DT) if want_results else None
        sec = self.lib.nnsp_oracle_cascade_batch_run(self._model_array(models), _p(seq_a), len(seq_a), _p(par),
                                                     S, _p(pcm2d), n, T, _p(res), n_threads)
        return sec, res


class RefLib:
    """The compiled, unmodified reference. NOT re-entrant: one stream at a time per process.

    dropin=True loads oracle/_ref/libnnsp_dropin.so instead: the reference's controller and model tables
    compiled against include/nnsp_compat and linked to libnnsp_b200.so, i.e. the same driver code on top of
    the CUDA engine's legacy entry points (needs a GPU to run)."""

    def __init__(self, acc32=False, dropin=False):
        name = "libnnsp_dropin.so" if dropin else "libnnsp_ref_acc%d.so" % (32 if acc32 else 64)
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        self.acc32 = bool(acc32)
        self.dropin = dropin
        assert dropin or L.ref_is_acc32() == int(self.acc32)
        L.ref_nnsp_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int16, C.c_int16] + [C.c_void_p] * 8
        if not dropin:
            L.ref_feature_stages.argtypes = [C.c_void_p] * 6
            L.ref_table.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.ref_cascade_run.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7

    @staticmethod
    def available(acc32=False, dropin=False):
        name = "libnnsp_dropin.so" if dropin else "libnnsp_ref_acc%d.so" % (32 if acc32 else 64)
        return os.path.exists(os.path.join(HERE, "_ref", name))

    def strides(self, nn_id):
        a, h, n = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_strides(nn_id, C.byref(a), C.byref(h), C.byref(n))
        return a.value, h.value, n.value

    def table(self, name):
        p, eb = C.c_void_p(), C.c_int()
        n = self.lib.ref_table(name.encode(), C.byref(p), C.byref(eb))
        dt = np.int16 if eb.value == 2 else np.int32
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16 if eb.value == 2 else C.c_int32)), (n,)).astype(dt)

    def feature_stages(self, win480):
        win480 = np.ascontiguousarray(win480, np.int16)
        out = dict(fft_in=np.zeros(512, np.int32), spec=np.zeros(514, np.int32),
                   pspec=np.zeros(257, np.int32), mel=np.zeros(40, np.int32), logmel=np.zeros(40, np.int32))
        self.lib.ref_feature_stages(_p(win480), _p(out["fft_in"]), _p(out["spec"]), _p(out["pspec"]),
                                    _p(out["mel"]), _p(out["logmel"]))
        return out

    def nnsp_run(self, nn_id, pcm, thresh_prob=16383, th_count=4, reset=True, taps=True):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = len(pcm) // 160
        a_s, h_s, n_o = self.strides(nn_id)
        res = np.zeros(T, RESULT_DT)
        tp = Taps(T, a_s, h_s, n_o) if taps else None
        a = [_p(getattr(tp, n)) for n in tp.names()] if taps else [None] * 7
        rc = self.lib.ref_nnsp_run(nn_id, int(reset), _p(pcm), T, thresh_prob, th_count, _p(res), *a)
        if rc:
            raise RuntimeError("ref_nnsp_run rc=%d" % rc)
        return res, tp

    def net_eval(self, nn_id, x, h, c):
        a_s, h_s, n_o = self.strides(nn_id)
        x = np.ascontiguousarray(x, np.int16)
        h = np.ascontiguousarray(h, np.int16).copy()
        c = np.ascontiguousarray(c, np.int32).copy()
        act = np.zeros(a_s, np.int16)
        logits = np.zeros(n_o, np.int32)
        self.lib.ref_net_eval.argtypes = [C.c_int] + [C.c_void_p] * 5
        rc = self.lib.ref_net_eval(nn_id, _p(x), _p(h), _p(c), _p(act), _p(logits))
        if rc:
            raise RuntimeError("ref_net_eval rc=%d" % rc)
        return act, logits, h, c

    def cascade_run(self, pcm, seq=(VAD, KWS, S2I), params=None, reset=True, taps=True):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = len(pcm) // 160
        seq_a = np.asarray(seq, np.int32)
        par = None if params is None else np.ascontiguousarray(params, np.int16)
        res = np.zeros(T, CASCADE_RESULT_DT)
        tp = Taps(T, 1, 128, 1) if taps else None
        valid = np.zeros(T, np.int8)
        a = [_p(tp.logmel), _p(tp.feat), _p(tp.h), _p(tp.c), _p(tp.post)] if taps else [None] * 5
        rc = self.lib.ref_cascade_run(int(reset), _p(seq_a), len(seq_a), _p(par), _p(pcm), T, _p(res), *a, _p(valid))
        if rc:
            raise RuntimeError("ref_cascade_run rc=%d" % rc)
        return res, tp, valid

    # -- synthetic models: a NeuralNetClass built over an NNSPM1 container (ref_glue.c ref_net_eval_blob) ------------
    def net_eval_blob(self, blob, acc32, x, h, c, act_stride, n_out):
        x = np.ascontiguousarray(x, np.int16)
        h = np.ascontiguousarray(h, np.int16).copy()
        c = np.ascontiguousarray(c, np.int32).copy()
        hh, cc = np.zeros(1024, np.int16), np.zeros(1024, np.int32)     # the glue reads / writes whole state rows
        hh[: len(h)] = h
        cc[: len(c)] = c
        act = np.zeros(max(act_stride, 1) + 512, np.int16)
        logits = np.zeros(n_out + 512, np.int32)
        buf = C.create_string_buffer(bytes(blob), len(blob))
        self.lib.ref_net_eval_blob.argtypes = [C.c_void_p, C.c_longlong, C.c_int] + [C.c_void_p] * 5
        rc = self.lib.ref_net_eval_blob(buf, len(blob), int(bool(acc32)), _p(x), _p(hh), _p(cc), _p(act), _p(logits))
        if rc:
            raise RuntimeError("ref_net_eval_blob rc=%d" % rc)
        return act[:act_stride], logits[:n_out], hh[: len(h)], cc[: len(c)]

    # -- many streams through the one controller, per-stream state swapped in and out (CPU arms of bench.py) ---------
    def cascade_state_bytes(self):
        self.lib.ref_cascade_state_bytes.restype = C.c_longlong
        return int(self.lib.ref_cascade_state_bytes())

    def cascade_batch(self, pcm2d, states, fresh, seq=(VAD, KWS, S2I), params=None, want_results=False, stage_frames=None):
        """pcm2d int16 [n, T*160]; states uint8 [n, cascade_state_bytes()] (kept by the caller between calls)."""
        pcm2d = np.ascontiguousarray(pcm2d, np.int16)
        n, ns = pcm2d.shape
        T = ns // 160
        seq_a = np.asarray(seq, np.int32)
        par = None if params is None else np.ascontiguousarray(params, np.int16)
        res = np.zeros((n, T), CASCADE_RESULT_DT) if want_results else None
        self.lib.ref_cascade_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.lib.ref_cascade_batch(int(bool(fresh)), _p(seq_a), len(seq_a), _p(par), n, _p(pcm2d), ns, T, _p(states),
                                        _p(res), _p(stage_frames))
        if rc:
            raise RuntimeError("ref_cascade_batch rc=%d" % rc)
        return res


class RefS2ICtrl:
    """oracle/_ref/libnnsp_ref_s2ictrl.so: the reference's S2I-only controller (evb/src/s2iCntrlClass.c), unmodified."""

    PATH = os.path.join(HERE, "_ref", "libnnsp_ref_s2ictrl.so")

    def __init__(self):
        if not os.path.exists(self.PATH):
            raise FileNotFoundError(self.PATH)
        self.lib = C.CDLL(self.PATH)
        self.lib.ref_s2ictrl_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def run(self, pcm, reset=1, thresh_prob=-1, thresh_cnts=-1):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = len(pcm) // 160
        res = np.zeros(T, CASCADE_RESULT_DT)
        rc = self.lib.ref_s2ictrl_run(int(reset), thresh_prob, thresh_cnts, _p(pcm), T, _p(res))
        if rc:
            raise RuntimeError("ref_s2ictrl_run rc=%d" % rc)
        return res
