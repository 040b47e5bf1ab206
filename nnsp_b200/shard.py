"""Multi-GPU host logic: streams are independent, so the path shards by a static block partition of stream
ids over ranks (one process per GPU), with NO collective on the data path. Result records are gathered to
every rank (or only used locally) over torch.distributed when a caller wants one global array
(SURVEY.md section 8e)."""
import numpy as np


def stream_range(total_streams, world_size, rank):
    """[first, last) stream ids owned by `rank`: contiguous blocks, sizes differ by at most one."""
    if not (0 <= rank < world_size) or total_streams < 0:
        raise ValueError("bad partition request")
    first = total_streams * rank // world_size
    last = total_streams * (rank + 1) // world_size
    return first, last


def gather_results(local, total_streams, group=None):
    """All ranks contribute their [n_local, T] structured result block; every rank gets [total, T].

    Blocks may differ in size by one stream, so they are padded to the largest block for all_gather."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    T = local.shape[1]
    item = local.dtype.itemsize
    sizes = [stream_range(total_streams, world, r) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    buf = np.zeros((nmax, T * item), np.uint8)
    buf[: local.shape[0]] = np.frombuffer(np.ascontiguousarray(local).tobytes(), np.uint8).reshape(local.shape[0], T * item)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(buf).to(dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    full = np.empty((total_streams, T), local.dtype)
    for r, (a, b) in enumerate(sizes):
        blk = outs[r].cpu().numpy()[: b - a]
        full[a:b] = np.frombuffer(blk.tobytes(), local.dtype).reshape(b - a, T)
    return full


def _physical_index(device):
    import os
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [x.strip() for x in vis.split(",") if x.strip() != ""]
        if device < len(ids) and ids[device].isdigit():
            return int(ids[device])
    return device


def nvml_handle(device):
    """NVML handle of CUDA device `device`: by PCI bus id when the CUDA runtime is there to ask (exact under any
    CUDA_VISIBLE_DEVICES, UUID entries included), else by the index CUDA_VISIBLE_DEVICES implies."""
    import pynvml
    pynvml.nvmlInit()
    try:
        from .api import device_pci_bus_id
        return pynvml.nvmlDeviceGetHandleByPciBusId(device_pci_bus_id(device).encode())
    except Exception:
        return pynvml.nvmlDeviceGetHandleByIndex(_physical_index(device))


def local_cpus(device):
    """CPUs on the NUMA node the GPU's PCIe root port hangs off (NVML), restricted to the ones this process may use;
    empty when NVML cannot tell."""
    import os
    try:
        import pynvml
        h = nvml_handle(device)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 1, 64) + 63) // 64)
    except Exception:
        return set()
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
    return cpus & os.sched_getaffinity(0)


def bind_host_to_device(device):
    """One process per GPU: keep the rank's host thread -- and with it the first-touch placement of the pinned PCM and
    result buffers it allocates next -- on the CPU socket the GPU is attached to, so that the H2D stream of every rank
    reads local DRAM instead of crossing the inter-socket link. Returns the CPU set bound to (empty = left alone)."""
    import os
    cpus = local_cpus(device)
    if cpus and cpus != os.sched_getaffinity(0):
        os.sched_setaffinity(0, cpus)
    return cpus
