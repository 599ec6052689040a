#!/usr/bin/env python
"""Benchmark of the log-mel front end (the one hot path; BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path (LogMelSpec.forward semantics: log-mel + batch scalar mean) over
one batch of synthetic audio.  At N = 1 the workload is BASELINE.json configs[1]: 64 x 30 s segments
(the ASR training chunk shape), 80 mel, 25 ms / 10 ms.  At N > 1 (torchrun, one rank per GPU) every
rank processes its own batch of the same shape (episode-sharded, weak scaling, no data-path
collective — the reference's mean is rank-local); value = frames of all ranks / max time.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (logmel_ws_kernel) vs the measured HBM copy bandwidth
  cpu_baseline  the oracle's fp32 port of the reference op sequence, timed on this box's host cores
  e2e           same metric through the public host-side API (HostPipeline) with pinned HOST buffers, copies inside the timing
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, SECONDS, SR, HOP, N_MELS = 64, 30, 16000, 160, 80
N_SAMPLES = SECONDS * SR
N_FRAMES = 1 + N_SAMPLES // HOP
FRAMES_PER_STEP = BATCH * N_FRAMES
ALGO_BYTES_PER_FRAME = 4 * HOP + 4 * N_MELS          # each sample read once, each feature written once
WORKLOAD = "configs[1]: batch of 64 x 30 s segments, 16 kHz mono, 80 mel, 25 ms/10 ms, batch scalar mean"
FALLBACK_HBM_GBS = 6650.0                            # B200_PROFILING.md fallback


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_rate(steps: int, warmup: int, batch_np=None, min_seconds: float = 0.0):
    """Times the oracle's fp32 port of the reference op sequence (oracle/logmel_oracle.py:logmel_port_f32,
    the restatement of tal/asr/models.py:36-53) on the host cores: `steps` passes, continued until `min_seconds` of
    timed work have accumulated (at most 300 passes).  Returns (frames/s, ms/pass, threads, best pass in s, passes)."""
    import torch
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import synth
    if batch_np is None:
        batch_np = synth.batch(2020, BATCH, N_SAMPLES)
    x = torch.from_numpy(batch_np)
    if os.environ.get("OMP_NUM_THREADS") in ("1", None) and "LOCAL_RANK" in os.environ:
        torch.set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1; the baseline may use every core
    threads = torch.get_num_threads()
    for _ in range(warmup):
        O.logmel_port_f32(x)
    times = []
    while len(times) < steps or (sum(times) < min_seconds and len(times) < 300):
        t0 = time.perf_counter()
        O.logmel_port_f32(x)
        times.append(time.perf_counter() - t0)
    mean_s = sum(times) / len(times)
    return FRAMES_PER_STEP / mean_s, mean_s * 1e3, threads, min(times), len(times)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    Python and cannot travel to the GPU box, and its arithmetic is the same torch.stft/matmul sequence)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    fps, ms, threads, best, _ = cpu_reference_rate(steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": "log-mel frames/sec", "value": fps, "unit": "frames/s",
        "realtime_factor": fps * 0.010, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"full 64 x 30 s batch per step, {steps} steps, torch CPU threads={threads}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tal_asrd_b200 import LogMelSpec, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the front end has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = _lib.load()
    mod = LogMelSpec(n_mels=N_MELS).to(dev)
    stream = torch.cuda.current_stream(dev)

    # synthetic inputs resident in HBM: NBUF distinct batches rotated so that a step never re-reads
    # what the previous two steps left in L2 (3 x 123 MB of input + 2 x 61 MB of output > 126 MB L2)
    NBUF = 3
    waves = []
    for i in range(NBUF):
        w = torch.empty(BATCH, N_SAMPLES, dtype=torch.float32, device=dev)
        _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, BATCH, N_SAMPLES, N_SAMPLES, 2020,
                                        (rank * NBUF + i) * BATCH, 0, stream.cuda_stream))
        waves.append(w)
    outs = [torch.empty(BATCH, N_FRAMES, N_MELS, dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()

    def step(i, norm="batch"):
        return mod.features(waves[i % NBUF], norm=norm, out=outs[i % 2])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        barrier()
        # keep the same load running a little longer so that NVML (ms-scale sampling) sees clocks under it
        t_end = time.time() + 0.5
        j = 0
        while time.time() < t_end:
            step(j); j += 1
            if j % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * FRAMES_PER_STEP / (ms_per_step * 1e-3)

    # dominant kernel alone (same launch geometry, un-normalised output): CUDA events on the launching stream
    for i in range(3):
        step(i, "none")
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        step(i, "none")
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    peaks, peak_kind = measured_peaks()
    achieved = FRAMES_PER_STEP * ALGO_BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            traffic = json.load(fh).get("logmel_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "logmel_ws_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step,
                "algorithmic_bytes_per_launch": FRAMES_PER_STEP * ALGO_BYTES_PER_FRAME}

    # end to end through the public API with HOST buffers (tal_asrd_b200.HostPipeline): every step's waveforms go
    # pinned host -> device, through the front end, and its features device -> pinned host, all inside the timed
    # region; consecutive steps overlap on three streams (H2D | transform | D2H), two device slots.  Two distinct
    # host batches alternate.  Also reported: the same step as ONE blocking call (forward_host: copy-in, transform,
    # copy-out back to back) and the pipeline fed with int16 PCM (the on-disk format).
    from tal_asrd_b200 import HostPipeline
    from tal_asrd_b200.hostpipe import bind_host_thread_to_gpu
    numa_cpus = bind_host_thread_to_gpu(local_rank) if world > 1 else None   # one rank per GPU: pin next to its own GPU
    host_in = [torch.empty(BATCH, N_SAMPLES, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        host_in[i].copy_(waves[i].cpu())
    host_out = [torch.empty(BATCH, N_FRAMES, N_MELS, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = max(3, min(args.steps, 20))

    def timed_pipeline(inputs):
        pipe = HostPipeline(mod, dev, depth=2)
        for i in range(3):
            pipe.submit(inputs[i % 2], host_out[i % 2])
        pipe.drain()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        p0.record(pipe.h2d)
        for i in range(e2e_steps):
            pipe.submit(inputs[i % 2], host_out[i % 2])
        p1.record(pipe.d2h)
        pipe.drain()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        tt = torch.tensor([p0.elapsed_time(p1), wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0].item()) / e2e_steps, float(tt[1].item()) / e2e_steps

    e2e_ms, e2e_wall_ms = timed_pipeline(host_in)

    mod.forward_host(host_in[0], host_out[0], device=dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        mod.forward_host(host_in[i % 2], host_out[i % 2], device=dev)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    blocking_ms = float(t.item()) / e2e_steps

    host_pcm = [(h * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory() for h in host_in]
    pcm_ms, _ = timed_pipeline(host_pcm)

    e2e = {"value": world * FRAMES_PER_STEP / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
           "wall_ms_per_step": e2e_wall_ms,
           "h2d_bytes_per_step": host_in[0].numel() * 4, "d2h_bytes_per_step": host_out[0].numel() * 4, "steps": e2e_steps,
           "api": "HostPipeline.submit per step (3 streams, 2 device slots), drain at the end; fp32 waveforms in, fp32 features out",
           "host_cpus_rank0": (f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)} CPUs local to the GPU)" if numa_cpus else "unbound"),
           "blocking_call": {"value": world * FRAMES_PER_STEP / (blocking_ms * 1e-3), "ms_per_step": blocking_ms,
                             "api": "LogMelSpec.forward_host (copy-in, transform, copy-out back to back)"},
           "pcm16_input": {"value": world * FRAMES_PER_STEP / (pcm_ms * 1e-3), "ms_per_step": pcm_ms,
                           "h2d_bytes_per_step": host_pcm[0].numel() * 2}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, ms, threads, best, passes = cpu_reference_rate(5, 2, host_in[0].numpy(), min_seconds=10.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "ms_per_step": ms,
               "best_ms_per_step": best * 1e3,
               "sample": f"the full 64 x 30 s batch (same samples as the GPU step), mean of {passes} passes "
                         f"(about 10 s of CPU work) after 2 warm-ups"}

    if rank == 0:
        line = {
            "metric": "log-mel frames/sec", "value": value, "unit": "frames/s", "realtime_factor": value * 0.010,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": FRAMES_PER_STEP,
                       "l2_policy": "3 rotating input batches (369 MB) + 2 output buffers (123 MB) > 126 MB L2",
                       "parallelism": f"episode-sharded x{world}, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": 2 * args.steps, "clocks": clocks.summary(),   # K1 + sub_scalar_flat_kernel per step
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


class _QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices) write to the
    process-level stdout from C, so file descriptor 1 is pointed at stderr for the duration of the run and the
    JSON line goes to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def emit(obj):
    line = json.dumps(obj)
    if _OUT is not None:
        _OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    with _QuietStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
        _OUT = None


if __name__ == "__main__":
    main()
