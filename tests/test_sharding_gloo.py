"""The N > 1 path on CPU: two processes over gloo, streams block-partitioned, no data-path collective, result
records gathered. The oracle stands in for the per-rank engine (this is a test of the host-side sharding logic;
the product ranks run the CUDA engine, see bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from common import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, T, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import nnsp_b200 as nb
    from nnsp_b200.shard import gather_results, stream_range
    from oracle.pyoracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = stream_range(total, world, rank)
    pcm = nb.synth_pcm(b - a, T, first_stream=a)          # each rank synthesises only its own streams
    O = Oracle()
    _, local = O.batch_run(O.model(1, False), pcm, n_threads=1, want_results=True)
    full = gather_results(local.view(nb.RESULT_DT), total)
    dist.barrier()
    if rank == 0:
        q.put(full.tobytes())
    dist.destroy_process_group()


def test_stream_range_is_a_partition(nb):
    from nnsp_b200.shard import stream_range
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            edges = [stream_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        stream_range(10, 2, 2)


@pytest.mark.timeout(180)
def test_two_ranks_gloo_equal_single_process(nb, oracle):
    import torch.multiprocessing as mp
    total, T, world = 13, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=150), nb.RESULT_DT).reshape(total, T)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pcm = nb.synth_pcm(total, T)
    _, ref = oracle.batch_run(oracle.model(1, False), pcm, want_results=True)
    assert (got == ref.view(nb.RESULT_DT)).all()


def test_host_binding_never_widens_or_empties_the_cpu_set(nb):
    """bind_host_to_device: the rank ends up on a non-empty subset of the CPUs it was allowed (the GPU's NVML CPU
    set where NVML answers, untouched otherwise -- this container has no GPU)."""
    from nnsp_b200.shard import bind_host_to_device, local_cpus
    before = os.sched_getaffinity(0)
    try:
        cpus = bind_host_to_device(0)
        after = os.sched_getaffinity(0)
        assert after and after <= before
        assert cpus == local_cpus(0) or not cpus
        assert after == (cpus if cpus else before)
    finally:
        os.sched_setaffinity(0, before)
