"""nnsp_b200_group_*: several devices (or several members on one device) behind one handle, host buffers in and out,
one host thread per member inside the library. Against a single handle over the same streams."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FILES = ("s2i.nnspm", "vad.nnspm", "kws_galaxy.nnspm")


def _devices(nb, n):
    have = nb.device_count()
    return [k % have for k in range(n)]              # on a one-GPU box: n members (n host threads) on device 0


@pytest.mark.parametrize("members", [2, 3])
def test_batch_group_equals_one_handle(nb, members):
    S, T = 1000, 90
    pcm = nb.synth_pcm(S, T, first_stream=2)
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + FILES[2], acc32=True)
    ref = nb.NNSPBatch(m, S)
    want = ref.exec_host(pcm)
    ref.close()
    g = nb.Group(m, S, _devices(nb, members))
    rng = g.ranges()
    assert len(rng) == members and rng[0][1] == 0 and sum(r[2] for r in rng) == S
    assert all(rng[k][1] + rng[k][2] == rng[k + 1][1] for k in range(members - 1))
    got = np.concatenate([g.exec_host(np.ascontiguousarray(pcm[:, :33 * 160])), g.exec_host(np.ascontiguousarray(pcm[:, 33 * 160:]))], axis=1)
    assert (got == want).all()
    g.reset()                                        # NNSPClass_reset on every member: the same audio again, from the start
    again = g.exec_host(pcm)
    ref = nb.NNSPBatch(m, S)
    ref.exec_host(pcm); ref.reset()
    assert (again == ref.exec_host(pcm)).all()
    ref.close(); g.close()


def test_cascade_group_async_calls(nb):
    S, n, calls = 900, 50, 5
    pcm = nb.synth_pcm(S, n * calls, first_stream=21)
    models = [nb.Model.from_blob(nb.MODEL_DIR + "/" + f) for f in FILES]
    params = dict(thresh_timeout_kws=60, thresh_timeout_s2i=40)
    ref = nb.Cascade(models, S, params=params)
    want = ref.exec_host(pcm)
    ref.close()
    g = nb.Group(models, S, _devices(nb, 2), params=params)
    pin = [nb.PinnedArray((S, n * 160), np.int16) for _ in range(2)]
    pres = [nb.PinnedArray((S, n), nb.CASCADE_RESULT_DT) for _ in range(2)]
    got, prev = [], None
    for k in range(calls):
        pin[k & 1].array[...] = pcm[:, k * n * 160:(k + 1) * n * 160]
        tk = g.exec_host_async(pin[k & 1].array, pres[k & 1].array)
        if prev is not None:
            g.wait(prev)
            got.append(pres[(k - 1) & 1].array.copy())
        prev = tk
    g.wait(prev)
    got.append(pres[(calls - 1) & 1].array.copy())
    got = np.concatenate(got, axis=1)
    for f in want.dtype.names:
        assert (got[f] == want[f]).all(), f
    for x in pin + pres:
        x.free()
    g.close()


def test_cascade_group_per_stream_params(nb):
    """nnsp_b200_group_set_stream_params: a range of streams that straddles the members gets its own time-outs after the
    first call; the group must then behave like one handle that was given the same parameter sets"""
    S, n = 300, 120
    pcm = nb.synth_pcm(S, 2 * n, first_stream=2)
    models = [nb.Model.from_blob(nb.MODEL_DIR + "/" + f) for f in FILES]
    ref = nb.Cascade(models, S)
    par = np.tile(ref.params_array(), (170, 1))
    names = [x for x, _ in nb.capi.CascadeParams._fields_]
    par[:, names.index("thresh_timeout_kws")] = 25 + np.arange(170) % 40
    par[:, names.index("thresh_cnts_vad")] = 1 + np.arange(170) % 5
    w1 = ref.exec_host(pcm[:, :n * 160].copy())
    ref.set_stream_params(60, par)
    w2 = ref.exec_host(pcm[:, n * 160:].copy())
    ref.close()
    g = nb.Group(models, S, _devices(nb, 3))
    g1 = g.exec_host(pcm[:, :n * 160].copy())
    g.set_stream_params(60, par)                          # members own [0,100), [100,200), [200,300): all three are touched
    g2 = g.exec_host(pcm[:, n * 160:].copy())
    g.close()
    for f in w1.dtype.names:
        assert (g1[f] == w1[f]).all() and (g2[f] == w2[f]).all(), f
    assert not all((w2[f][60:230] == nb_default_second_call(nb, models, pcm, n)[f][60:230]).all() for f in w2.dtype.names)


def nb_default_second_call(nb, models, pcm, n):
    c = nb.Cascade(models, pcm.shape[0])
    c.exec_host(pcm[:, :n * 160].copy())
    r = c.exec_host(pcm[:, n * 160:].copy())
    c.close()
    return r


def test_group_over_every_device_of_the_box(nb):
    n = nb.device_count()
    if n < 2:
        pytest.skip("one device on this box (the multi-member logic is covered above)")
    S, T = 64 * n + 5, 40
    pcm = nb.synth_pcm(S, T, first_stream=7)
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + FILES[1])
    ref = nb.NNSPBatch(m, S)
    want = ref.exec_host(pcm)
    ref.close()
    g = nb.Group(m, S, list(range(n)))
    assert [r[0] for r in g.ranges()] == list(range(n))
    assert (g.exec_host(pcm) == want).all()
    g.close()
