"""GPU parity of the batched nnCntrlClass cascade (evb/src/nnCntrlClass.c:152-272) against the oracle:
stage id, position, detection, outputs and timeout counter of every frame, plus the live instance's
feature row / LSTM state / NNSPClass scalars, through stage changes, look-back and timeouts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MODEL_FILE = {0: "s2i.nnspm", 1: "vad.nnspm", 2: "kws_galaxy.nnspm"}


def _models(nb, acc32=False):
    return [nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[i], acc32=acc32) for i in range(3)]


def _oracle_models(oracle, acc32=False):
    return [oracle.model(i, acc32) for i in range(3)]


def _check(res_gpu, taps_gpu, s, res_or, taps_or, valid):
    for f in ("stage_id", "pos_after", "detected", "outputs", "cnt_timeout"):
        if not (res_gpu[s][f] == res_or[f]).all():
            t = int(np.nonzero((res_gpu[s][f] != res_or[f]).reshape(len(res_or), -1).any(axis=1))[0][0])
            raise AssertionError("stream %d frame %d: %s gpu %s oracle %s (stage %d)" % (s, t, f, res_gpu[s][f][t], res_or[f][t], res_or["stage_id"][t]))
    if taps_gpu is None:
        return
    for g, o in (("logmel", "logmel"), ("feat", "feat"), ("hstate", "h"), ("cstate", "c"), ("post", "post")):
        a, b = taps_gpu[g][s], getattr(taps_or, o)
        if not (a == b).all():
            t = int(np.nonzero((a != b).any(axis=1))[0][0])
            raise AssertionError("stream %d frame %d: tap %s differs (stage %d valid %d): gpu %s oracle %s" %
                                 (s, t, g, res_or["stage_id"][t], valid[t], a[t][:6], b[t][:6]))


def test_cascade_default_sequence_every_frame(nb, oracle):
    S, T = 20, 2600                 # long enough for detections, look-back and the 999-frame timeouts
    pcm = nb.synth_pcm(S, T, first_stream=4)
    c = nb.Cascade(_models(nb), S)
    res, taps = c.exec(pcm, taps=True)
    om = _oracle_models(oracle)
    stages, transitions = set(), 0
    for s in range(S):
        r, tp, valid = oracle.cascade_run(om, pcm[s])
        _check(res, taps, s, r, tp, valid)
        stages |= set(np.unique(r["stage_id"]).tolist())
        transitions += int((np.diff(r["pos_after"].astype(int)) != 0).sum())
    assert stages == {0, 1, 2} and transitions > 10      # the workload really exercises the controller
    c.close()


def test_cascade_chunked_equals_one_shot(nb, oracle):
    S, T = 12, 700
    pcm = nb.synth_pcm(S, T, first_stream=20)
    models = _models(nb)
    c = nb.Cascade(models, S)
    full = c.exec(pcm)
    c.close()
    c = nb.Cascade(models, S)
    parts, t = [], 0
    for n in (1, 1, 3, 80, 1, 2, 79, 200, 33, 300):
        parts.append(c.exec(pcm[:, t * 160:(t + n) * 160]))
        t += n
    assert t == T
    chunked = np.concatenate(parts, axis=1)
    for f in full.dtype.names:
        assert (chunked[f] == full[f]).all(), f
    c.close()


@pytest.mark.parametrize("seq,params", [
    ((1, 0), dict(frs_vbufBk_s2i=5, thresh_timeout_s2i=60, thresh_cnts_vad=2)),
    ((0,), dict(frs_vbufBk_s2i=0, thresh_timeout_s2i=37, thresh_cnts_s2i=1)),
    ((1, 2, 0), dict(frs_vbufBk_kws=99, frs_vbufBk_s2i=1, thresh_timeout_kws=45, thresh_timeout_s2i=30, thresh_prob_kws=100, thresh_cnts_kws=2)),
])
def test_cascade_other_sequences_and_params(nb, oracle, seq, params):
    S, T = 16, 600
    pcm = nb.synth_pcm(S, T, first_stream=64)
    c = nb.Cascade(_models(nb), S, seq=seq, params=params)
    res, taps = c.exec(pcm, taps=True)
    om = _oracle_models(oracle)
    par = c.params_array()
    for s in range(S):
        r, tp, valid = oracle.cascade_run(om, pcm[s], seq=seq, params=par)
        _check(res, taps, s, r, tp, valid)
    c.close()


def test_cascade_reset_midstream_and_acc32(nb, oracle):
    S, T = 10, 500
    pcm = nb.synth_pcm(S, 2 * T, first_stream=128)
    c = nb.Cascade(_models(nb, acc32=True), S)
    c.exec(pcm[:, :T * 160])
    c.reset()                                       # nnCntrlClass_reset: position kept, stale context rows kept
    res, taps = c.exec(pcm[:, T * 160:], taps=True)
    om = _oracle_models(oracle, acc32=True)
    for s in range(S):
        st = oracle.lib.nnsp_oracle_cascade_new()
        oracle.cascade_run(om, pcm[s, :T * 160], state=st, reset=1, taps=False)
        r, tp, valid = oracle.cascade_run(om, pcm[s, T * 160:], state=st, reset=2)
        oracle.lib.nnsp_oracle_cascade_free(st)
        _check(res, taps, s, r, tp, valid)
    c.close()


def test_cascade_exec_host(nb):
    S, T = 260, 120
    pcm = nb.synth_pcm(S, T, first_stream=3)
    models = _models(nb)
    c = nb.Cascade(models, S)
    dev = c.exec(pcm)
    c.close()
    c = nb.Cascade(models, S)
    host = c.exec_host(pcm)
    for f in dev.dtype.names:
        assert (dev[f] == host[f]).all(), f
    c.close()


@pytest.mark.parametrize("path", ["sorted", "sequential"])
def test_cascade_async_host_calls_double_buffered(nb, path):
    """nnsp_b200_cascade_exec_host_async / _wait_host with two pinned buffer pairs against one device-buffer call over the
    same audio; a blocking host call and a device call follow the asynchronous ones."""
    S, n, calls = 4100, 40, 6                          # >= 4096 streams: 8 slices over the 3 pipeline streams
    T = n * (calls + 2)
    pcm = nb.synth_pcm(S, T, first_stream=11)
    models = _models(nb)
    ref = nb.Cascade(models, S)
    ref.set_path(path)
    want = ref.exec(pcm)
    ref.close()
    c = nb.Cascade(models, S)
    c.set_path(path)
    pin = [nb.PinnedArray((S, n * 160), np.int16) for _ in range(2)]
    pres = [nb.PinnedArray((S, n), nb.CASCADE_RESULT_DT) for _ in range(2)]
    got, prev = [], None
    for k in range(calls):
        pin[k & 1].array[...] = pcm[:, k * n * 160:(k + 1) * n * 160]
        tk = c.exec_host_async(pin[k & 1].array, pres[k & 1].array)
        if prev is not None:
            c.wait_host(prev)
            got.append(pres[(k - 1) & 1].array.copy())
        prev = tk
    c.wait_host(prev)
    got.append(pres[(calls - 1) & 1].array.copy())
    got.append(c.exec_host(pcm[:, calls * n * 160:(calls + 1) * n * 160].copy()))
    got.append(c.exec(pcm[:, (calls + 1) * n * 160:].copy()))
    got = np.concatenate(got, axis=1)
    for f in want.dtype.names:
        assert (want[f] == got[f]).all(), f
    for x in pin + pres:
        x.free()
    c.close()


@pytest.mark.parametrize("seq,params,chunks", [
    ((1, 2, 0), None, (100, 100, 37, 1, 2, 160, 300, 100, 100, 100, 100, 100, 100, 100, 100, 100, 100)),
    ((1, 2, 0), dict(frs_vbufBk_kws=99, frs_vbufBk_s2i=1, thresh_timeout_kws=45, thresh_timeout_s2i=30, thresh_prob_kws=100, thresh_cnts_kws=2),
     (64, 3, 100, 1, 1, 31, 100, 100, 200)),
    ((1, 0), dict(frs_vbufBk_s2i=5, thresh_timeout_s2i=60, thresh_cnts_vad=2), (100, 100, 100, 50, 7, 143, 100)),
    ((0,), dict(frs_vbufBk_s2i=0, thresh_timeout_s2i=37, thresh_cnts_s2i=1), (100, 99, 101, 100)),
])
def test_cascade_stage_sorted_pass(nb, oracle, seq, params, chunks):
    """The stage-sorted pass (scan-split kernels per (model, phase) group, controller walk, replay after a stage
    change) against the oracle on every frame, any chunking; the last chunk runs on the sequential kernel with
    all taps, so the state the sorted pass leaves behind (context, LSTM, counters, stale rows) is checked too."""
    S = 75                                             # 4 full tiles + a partial one, spread over the groups
    T = sum(chunks)
    pcm = nb.synth_pcm(S, T, first_stream=17)
    c = nb.Cascade(_models(nb), S, seq=seq, params=params)
    c.set_path("sorted")
    parts, t = [], 0
    for n in chunks[:-1]:
        parts.append(c.exec(pcm[:, t * 160:(t + n) * 160]))
        t += n
    c.set_path("sequential")
    last, taps = c.exec(pcm[:, t * 160:], taps=True)
    res = np.concatenate(parts + [last], axis=1)
    om = _oracle_models(oracle)
    par = c.params_array()
    transitions = 0
    for s in range(S):
        r, tp, valid = oracle.cascade_run(om, pcm[s], seq=seq, params=par)
        _check(res, None, s, r, None, None)
        for g, o in (("feat", "feat"), ("hstate", "h"), ("cstate", "c"), ("post", "post")):
            assert (taps[g][s] == getattr(tp, o)[t:]).all(), "state after the sorted pass: tap %s stream %d" % (g, s)
        transitions += int((np.diff(r["pos_after"][:t].astype(int)) != 0).sum())
    if len(seq) > 1:
        assert transitions > 5                         # stage changes happened inside sorted-pass chunks
    c.close()


def test_cascade_sorted_exec_host_equals_sequential(nb):
    S, T = 530, 120
    pcm = nb.synth_pcm(S, 2 * T, first_stream=9)
    models = _models(nb)
    out = {}
    for path in ("sequential", "sorted"):
        c = nb.Cascade(models, S)
        c.set_path(path)
        a = c.exec_host(pcm[:, :T * 160].copy())
        b = c.exec_host(pcm[:, T * 160:].copy())
        out[path] = np.concatenate([a, b], axis=1)
        c.close()
    for f in out["sorted"].dtype.names:
        assert (out["sorted"][f] == out["sequential"][f]).all(), f


def test_cascade_back_to_back_calls_are_pipelined_correctly(nb, oracle):
    """Cascade calls issued without a sync in between (front end of call N+1 overlaps the controller / network work of
    call N; log-mel rows and the PCM history are double buffered, the replay reads the previous call's PCM)."""
    S = 60
    lens = [50, 50, 3, 47, 100, 50, 1, 49, 50, 100]
    T = sum(lens)
    pcm = nb.synth_pcm(S, T, first_stream=33)
    params = dict(frs_vbufBk_kws=99, frs_vbufBk_s2i=1, thresh_timeout_kws=45, thresh_timeout_s2i=30, thresh_prob_kws=100, thresh_cnts_kws=2)
    c = nb.Cascade(_models(nb), S, params=params)
    c.set_path("sorted")
    d_pcm, d_res, t = [], [], 0
    for n in lens:
        d_pcm.append(nb.DeviceArray.from_host(pcm[:, t * 160:(t + n) * 160]))
        d_res.append(nb.DeviceArray((S, n), nb.CASCADE_RESULT_DT))
        t += n
    for k, n in enumerate(lens):
        c.exec_device(d_pcm[k], n * 160, n, d_res[k])
    c.sync()
    res = np.concatenate([r.to_host() for r in d_res], axis=1)
    om = _oracle_models(oracle)
    par = c.params_array()
    for s in range(S):
        r, tp, valid = oracle.cascade_run(om, pcm[s], params=par)
        _check(res, None, s, r, None, None)
    for x in d_pcm + d_res:
        x.free()
    c.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_cascade_many_pipelined_calls_equal_sequential_one_shot(nb, seed):
    """30 back-to-back cascade calls of random lengths on 700 streams (stage-sorted pass, pipelined) against one call
    of the sequential kernel over the same 400 frames."""
    rng = np.random.default_rng(seed)
    S, T = 700, 400
    cuts = np.sort(rng.choice(np.arange(1, T), 29, replace=False))
    lens = np.diff(np.concatenate([[0], cuts, [T]])).tolist()
    pcm = nb.synth_pcm(S, T, first_stream=seed * 1000)
    params = dict(frs_vbufBk_kws=40, frs_vbufBk_s2i=3, thresh_timeout_kws=70, thresh_timeout_s2i=50, thresh_prob_kws=100, thresh_cnts_kws=2)
    models = _models(nb)
    ref = nb.Cascade(models, S, params=params)
    ref.set_path("sequential")
    want = ref.exec(pcm)
    ref.close()
    c = nb.Cascade(models, S, params=params)
    c.set_path("sorted")
    d_pcm, d_res, t = [], [], 0
    for n in lens:
        d_pcm.append(nb.DeviceArray.from_host(pcm[:, t * 160:(t + n) * 160]))
        d_res.append(nb.DeviceArray((S, n), nb.CASCADE_RESULT_DT))
        t += n
    for k, n in enumerate(lens):
        c.exec_device(d_pcm[k], n * 160, n, d_res[k])
    c.sync()
    got = np.concatenate([r.to_host() for r in d_res], axis=1)
    for f in got.dtype.names:
        assert (got[f] == want[f]).all(), f
    assert len(np.unique(want["stage_id"])) >= 2 and (np.diff(want["pos_after"].astype(int), axis=1) != 0).sum() > 100
    for x in d_pcm + d_res:
        x.free()
    c.close()


def test_s2i_only_controller_on_the_gpu(nb):
    """evb/src/s2iCntrlClass.c (unmodified reference, oracle/_ref/libnnsp_ref_s2ictrl.so) == nnsp_b200_cascade with
    seq = {s2i}, frs_vbufBk_s2i = 0 -- both cascade paths, detections and the resets that follow them included."""
    from oracle.pyoracle import RefS2ICtrl
    if not RefS2ICtrl.available():
        pytest.skip("oracle/_ref/libnnsp_ref_s2ictrl.so not built")
    R = RefS2ICtrl()
    x = np.concatenate([nb.synth_pcm(12, 700, first_stream=17), nb.synth_pcm(1, 700, first_stream=5)])
    models = _models(nb)
    for th_cnt in (4, 1):
        want = [R.run(x[s], reset=1, thresh_cnts=th_cnt) for s in range(len(x))]
        assert sum(int(w["detected"].sum()) for w in want) > 0
        for path in ("sorted", "sequential"):
            c = nb.Cascade(models, len(x), seq=(0,), params=dict(frs_vbufBk_s2i=0, thresh_cnts_s2i=th_cnt))
            c.set_path(path)
            got = np.concatenate([c.exec(x[:, :300 * 160]), c.exec(x[:, 300 * 160:])], axis=1)
            c.close()
            for s in range(len(x)):
                for f in ("stage_id", "pos_after", "detected", "outputs"):
                    assert (got[s][f] == want[s][f]).all(), (path, th_cnt, s, f)


@pytest.mark.parametrize("path", ["sorted", "sequential"])
def test_per_stream_params_changed_at_run_time(nb, oracle, path):
    """ParamCntrlClass is a member of every nnCntrlClass instance (nnCntrlClass.h:12-29): each stream gets its own
    thresholds, counts and time-outs (nnsp_b200_cascade_set_stream_params), and a third of the streams get new ones between
    two calls -- the oracle runs every stream as its own instance with exactly those parameter sets."""
    S, T1, T2 = 40, 260, 340
    rng = np.random.default_rng(11)
    pcm = nb.synth_pcm(S, T1 + T2, first_stream=700)
    c = nb.Cascade(_models(nb), S)
    c.set_path(path)
    base = c.params_array()
    names = [n for n, _ in nb.capi.CascadeParams._fields_]

    def draw():
        p = base.copy()
        for n, lo, hi in [("thresh_timeout_kws", 20, 120), ("thresh_timeout_s2i", 15, 90), ("thresh_cnts_vad", 1, 6), ("thresh_cnts_kws", 1, 5),
                          ("thresh_cnts_s2i", 1, 5), ("thresh_prob_vad", 2000, 30000), ("thresh_prob_kws", 100, 30000), ("thresh_prob_s2i", 100, 30000)]:
            p[names.index(n)] = rng.integers(lo, hi + 1)
        return p

    par1 = np.stack([draw() for _ in range(S)])
    c.set_stream_params(0, par1)
    r1 = c.exec(pcm[:, :T1 * 160])
    par2 = par1.copy()
    first, n = 13, S // 3
    par2[first:first + n] = np.stack([draw() for _ in range(n)])
    c.set_stream_params(first, par2[first:first + n])
    r2, taps = c.exec(pcm[:, T1 * 160:], taps=True)
    om = _oracle_models(oracle)
    for s in range(S):
        st = oracle.lib.nnsp_oracle_cascade_new()
        o1, _, _ = oracle.cascade_run(om, pcm[s, :T1 * 160], params=par1[s], state=st, reset=1, taps=False)
        o2, tp, valid = oracle.cascade_run(om, pcm[s, T1 * 160:], params=par2[s], state=st, reset=0)
        oracle.lib.nnsp_oracle_cascade_free(st)
        for f in o1.dtype.names:
            assert (r1[s][f] == o1[f]).all(), "first call, stream %d, field %s" % (s, f)
        _check(r2, taps, s, o2, tp, valid)
    bad = par1[:1].copy()
    bad[0, names.index("frs_vbufBk_kws")] += 1               # the look-back depth sizes the history: per handle only
    with pytest.raises(nb.NnspError):
        c.set_stream_params(0, bad)
    c.close()
