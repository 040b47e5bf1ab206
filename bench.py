#!/usr/bin/env python
"""bench.py -- throughput of the ns-nnsp streaming hot path on B200, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the VAD model (def_nn1_vad, 64-bit accumulators) over 4,096
independent synthetic 16 kHz streams per GPU; one step = 100 frames (1.0 s of audio) of every stream
through FeatureClass -> NeuralNetClass -> NNSPClass post-processing.

One JSON line on stdout (rank 0):
  value    audio-seconds per second, PCM already resident in HBM, CUDA-event timed on the engine's stream
  e2e      the same metric through the C ABI call with HOST (pinned) buffers: H2D + kernels + D2H inside
  roofline dominant kernel (feat_kernel): algorithmic bytes per launch / CUDA-event duration vs measured HBM peak
  cpu_baseline  the reference's own C (oracle/_ref, unmodified sources, gcc -O2) on this host's cores
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = 160
STREAMS_PER_GPU = 4096
FRAMES_PER_STEP = 100
REF_SAMPLE_STREAMS = STREAMS_PER_GPU                 # CPU arms: every stream of one step (~0.5 s on 16 cores per pass)
MODEL_ID = 1                                     # VAD
ALGO_BYTES_PER_STREAM_FRAME = 2608               # SURVEY.md section 8(d), VAD, one-frame-per-launch design
ALGO_INTOPS_PER_FRAME_FEATURE = 17700            # SURVEY.md section 8(d), feature stage
AUDIO_S_PER_FRAME = FRAME / 16000.0


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on the bench workload, from the
    committed ncu --set full capture (profiles/traffic.json); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)[kernel]
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


def measured_counts(kernel):
    """warp instructions executed by one launch of `kernel` on the bench workload (committed ncu capture), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process (a sample every 2 ms, so even a
    13 ms region gets several), nvidia-smi as the fall-back when pynvml is not importable."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.sm, self.mx, self.reasons, self._stop_evt = device, [], [], set(), threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            from nnsp_b200.shard import nvml_handle
            self.nvml, self.handle = pynvml, nvml_handle(device)      # the GPU the kernels run on, by PCI bus id
            self.mx.append(int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if device < len(ids) and ids[device].strip().isdigit():
                return int(ids[device])
        return device

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        if not self.mx:                                  # a property of the board: asked once, the loop stays short
            self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.handle))
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self._physical_index(self.device)), "--query-gpu=" + self.FIELDS,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6 and parts[0].isdigit():
            self.sm.append(int(parts[0]))
            if parts[1].isdigit():
                self.mx.append(int(parts[1]))
            for i in range(4):
                if parts[2 + i].lower().startswith("active"):
                    self.reasons.add(self.NAMES[i])

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.0005 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own C implementation on the host cores
# ---------------------------------------------------------------------------------------------
_REF_PCM = None          # set before the pool forks: the workers inherit the sample instead of receiving it through a pipe


def _ref_worker(args):
    kind, nn_id, lo, hi, T = args
    pcm = _REF_PCM[lo:hi]
    sys.path.insert(0, ROOT)
    from oracle import pyoracle
    t0 = time.perf_counter()
    if kind == "reference":
        R = pyoracle.RefLib(False)
        for row in pcm:                                   # single-instance library: stream after stream
            R.nnsp_run(nn_id, row, taps=False)
    else:
        O = pyoracle.Oracle()
        O.batch_run(O.model(nn_id, False), pcm, n_threads=1)
    return time.perf_counter() - t0


def cpu_reference_throughput(n, T, cores, pool):
    """audio-s/s of the reference C over the first `n` streams of _REF_PCM ([*, T*160]) split across `cores` forked
    processes (processes, not threads: the reference keeps global scratch)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.RefLib.available(False) else "port"
    chunks = [(n * k // cores, n * (k + 1) // cores) for k in range(cores)]
    chunks = [c for c in chunks if c[1] > c[0]]
    t0 = time.perf_counter()
    pool.map(_ref_worker, [(kind, MODEL_ID, lo, hi, T) for lo, hi in chunks], chunksize=1)
    dt = time.perf_counter() - t0
    return n * T * AUDIO_S_PER_FRAME / dt, kind, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from nnsp_b200.synth import synth_pcm
    cores = os.cpu_count() or 1
    T = FRAMES_PER_STEP
    global _REF_PCM
    n = REF_SAMPLE_STREAMS
    _REF_PCM = synth_pcm(n, T)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_throughput(n, T, cores, pool)
        t0 = time.perf_counter()
        kind = "port"
        for _ in range(args.steps):
            _, kind, _ = cpu_reference_throughput(n, T, cores, pool)
        dt = time.perf_counter() - t0
    value = args.steps * n * T * AUDIO_S_PER_FRAME / dt
    sample = "%d of the %d streams x %d frames per step, %d forked processes, oracle/%s" % (
        n, STREAMS_PER_GPU, T, cores, "_ref (unmodified reference C, gcc -O2)" if kind == "reference" else "nnsp_oracle.c port")
    line = {"impl": "reference", "metric": "audio-sec/sec (16 kHz streams, VAD, FeatureClass+NeuralNetClass+NNSPClass)",
            "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 activations x int8 weights, int64 accumulate", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, n_bound=0):
    return {"workload": "VAD model (def_nn1_vad, acc64) x %d streams per GPU x %d frames (%.1f s audio) per step; "
                        "FeatureClass -> NeuralNetClass -> NNSPClass post-proc, bit-exact vs reference C" %
                        (STREAMS_PER_GPU, FRAMES_PER_STEP, FRAMES_PER_STEP * AUDIO_S_PER_FRAME),
            "streams_per_gpu": STREAMS_PER_GPU, "frames_per_step": FRAMES_PER_STEP, "total_streams": STREAMS_PER_GPU * n_gpus,
            "partition": "streams block-partitioned over GPUs, no data-path collective",
            "cache": "two alternating %d MB PCM buffers per GPU (larger than the 126 MB L2)" %
                     (STREAMS_PER_GPU * FRAMES_PER_STEP * FRAME * 2 // 1000000),
            "host": "rank bound to the %d CPUs local to its GPU (NVML affinity)" % n_bound if n_bound else "rank not bound to a CPU set"}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import nnsp_b200 as nb
    from nnsp_b200.synth import synth_pcm

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    S, T = STREAMS_PER_GPU, FRAMES_PER_STEP

    # CPU baseline first (rank 0, N=1 only), before the GPU context makes fork() unsafe
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        global _REF_PCM
        n = REF_SAMPLE_STREAMS
        _REF_PCM = synth_pcm(n, T)
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_throughput(max(cores, 8), T, cores, pool)                      # warm the processes
            reps, t_total, val, kind = 0, 0.0, 0.0, "port"
            while (t_total < 1.5 or reps < 3) and reps < 8:                              # ~25 core-seconds on 16 cores
                v, kind, dt = cpu_reference_throughput(n, T, cores, pool)
                val, t_total, reps = max(val, v), t_total + dt, reps + 1
        _REF_PCM = None
        cpu_baseline = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind,
                        "sample": "%d of the %d streams x %d frames, best of %d passes, %d forked processes, %s" % (
                            n, S, T, reps, cores, "oracle/_ref = unmodified reference C (gcc -O2)" if kind == "reference" else "oracle/nnsp_oracle.c port")}

    # after the CPU baseline (its workers must keep every core): host thread and pinned buffers next to the GPU
    bound = set()
    if os.environ.get("NNSP_BENCH_BIND", "1") != "0":
        from nnsp_b200.shard import bind_host_to_device
        bound = bind_host_to_device(device)

    model = nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad.nnspm"), acc32=False)
    batch = nb.NNSPBatch(model, S, device=device)
    # two distinct PCM sets so consecutive steps never re-read L2-resident input
    host_pcm = [synth_pcm(S, T, first_stream=rank * S + k * 1000003) for k in range(2)]
    dev_pcm = [nb.DeviceArray.from_host(p, device) for p in host_pcm]
    dev_res = nb.DeviceArray((S, T), nb.RESULT_DT, device)
    ev0, ev1 = nb.Event(device), nb.Event(device)

    def barrier():
        batch.sync()
        if dist is not None:
            dist.barrier()
        batch.sync()

    def step(i):
        batch.exec_device(dev_pcm[i & 1], T * FRAME, T, dev_res)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(device)
    sampler.start()
    n0 = nb.kernel_launches()
    feat_ms = nn_ms = 0.0
    # calls are asynchronous and pipelined inside the engine (front end of step i+1 on one CUDA stream while the network
    # kernels of step i finish on another): the closing event is recorded once every stream of the handle has drained
    ev0.record(batch.stream)
    for i in range(args.steps):
        step(i)
    batch.sync()
    ev1.record(batch.stream)
    batch.sync()
    ms_total = ev0.elapsed_ms_to(ev1)
    launches = nb.kernel_launches() - n0
    clocks = sampler.stop()
    # per-kernel durations (CUDA events the engine records around its own launches on the same stream)
    km = []
    for i in range(min(args.steps, 5)):
        step(i)
        batch.sync()
        km.append(batch.last_kernel_ms())
    feat_ms = float(np.mean([k[0] for k in km]))
    nn_ms = float(np.mean([k[1] for k in km]))
    barrier()

    # end to end through the host-buffer entry point (pinned host memory -> H2D -> kernels -> D2H): the serving loop of
    # INTEGRATION.md -- two pinned PCM/result buffer pairs, call i queued while the results of call i-1 are consumed
    pin = [nb.PinnedArray((S, T * FRAME), np.int16) for _ in range(2)]
    for k in range(2):
        pin[k].array[...] = host_pcm[k]
    pin_res = [nb.PinnedArray((S, T), nb.RESULT_DT) for _ in range(2)]
    trig = [r.array["trigger"] for r in pin_res]

    def e2e_loop(n):
        fired, prev = 0, None
        for i in range(n):
            tk = batch.exec_host_async(pin[i & 1].array, pin_res[i & 1].array)
            if prev is not None:
                batch.wait_host(prev)
                fired += int(np.count_nonzero(trig[(i - 1) & 1]))        # the host reads the step's result
            prev = tk
        batch.wait_host(prev)
        return fired + int(np.count_nonzero(trig[(n - 1) & 1]))

    e2e_loop(3)
    barrier()
    sampler = ClockSampler(device)
    sampler.start()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_s = time.perf_counter() - t0
    clocks_e2e = sampler.stop()
    clocks["e2e"] = {k: clocks_e2e[k] for k in ("sm_mhz", "reasons", "samples")}
    clocks["reasons"] = sorted(set(clocks["reasons"]) | set(clocks_e2e["reasons"]))
    barrier()
    # the same with one blocking call per step (nothing in flight across calls)
    t0 = time.perf_counter()
    for i in range(args.steps):
        batch.exec_host(pin[i & 1].array, pin_res[i & 1].array)
    e2e_sync_s = time.perf_counter() - t0
    barrier()

    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        audio_s = world * S * T * AUDIO_S_PER_FRAME * args.steps
        value = audio_s / (ms_total * 1e-3)
        peak, peak_src = measured_peaks()
        frames_per_launch = S * T
        ach = ALGO_BYTES_PER_STREAM_FRAME * frames_per_launch / (feat_ms * 1e-3) / 1e9
        ipk = nb.int_peak(device)
        imad, mixed = ipk["imad"], ipk["mixed"]
        int_ach = ALGO_INTOPS_PER_FRAME_FEATURE * frames_per_launch / (feat_ms * 1e-3) / 1e9
        cnt = measured_counts("feat_kernel") or {}
        winst = cnt.get("warp_instructions_per_launch")
        sm_mhz = clocks.get("sm_mhz") or 1965
        issue_peak = 4.0 * 148 * sm_mhz * 1e6                      # one warp instruction per clock per SM sub-partition
        issue = None if not winst else {
            "kernel": "feat_kernel", "warp_instructions_per_launch": winst,
            "achieved_ginst_s": winst / (feat_ms * 1e-3) / 1e9, "peak_ginst_s": issue_peak / 1e9,
            "frac": winst / (feat_ms * 1e-3) / issue_peak,
            "pipe_fmaheavy_pct": cnt.get("pipe_fmaheavy_pct"), "pipe_alu_pct": cnt.get("pipe_alu_pct"),
            "note": "the binding resource of this integer kernel: issue slots and the FMA-heavy pipe (IMAD.WIDE = 4 cycles); "
                    "instruction and pipe counts from the committed ncu capture (profiles/traffic.json), duration measured live"}
        line = {
            "metric": "audio-sec/sec (16 kHz streams, VAD, FeatureClass+NeuralNetClass+NNSPClass)",
            "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 activations x int8 weights, int32/int64 accumulate (bit-exact integer path)",
            "data": "synthetic", "config": workload_config(world, len(bound)),
            "e2e": {"value": audio_s / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": S * T * FRAME * 2,
                    "d2h_bytes_per_step": S * T * 8, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "nnsp_b200_batch_exec_host_async + _wait_host, two pinned buffer pairs, results of every step read on the host",
                    "blocking_call_value": audio_s / world * 1.0 / e2e_sync_s if world == 1 else None},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "feat_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": measured_traffic("feat_kernel"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STREAM_FRAME * frames_per_launch,
                         "launch_ms": feat_ms,
                         "note": "achieved = SURVEY.md 8(d) whole-path algorithmic bytes (2608 B per stream-frame, VAD) / feat_kernel "
                                 "launch time; the kernel itself moves 154.8 MB per launch (traffic, = its own 320 B PCM in + 80 B feature "
                                 "row out per frame, no re-reads). The path is integer-issue bound, not HBM bound: see int_alu"},
            "int_alu": {"kernel": "feat_kernel", "achieved_gops": int_ach, "peak_gops_imad": imad,
                        "peak_gops_mixed": mixed, "peak_gops_imad_wide": ipk["imad_wide"], "peak_gops_idp2a": ipk["idp2a"], "frac_of_mixed_peak": int_ach / mixed if mixed else None,
                        "algorithmic_int_ops_per_frame": ALGO_INTOPS_PER_FRAME_FEATURE,
                        "peak_source": "self-measured nnsp_b200_int_peak (register-resident IMAD / IMAD+ALU chains)"},
            "issue": issue,
            "kernel_ms": {"feat_kernel": feat_ms, "network_kernels": nn_ms,
                          "network_path": "scan-split: seg_kernel<feat> + scan_kernel + seg_kernel<planes> + post_kernel + ctx_kernel"},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
