#!/usr/bin/env python
"""Benchmark of the log-mel front end (the one hot path; BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path (LogMelSpec.forward semantics: log-mel + batch scalar mean) over one batch of
synthetic audio.  At N = 1 the workload is BASELINE.json configs[1]: 64 x 30 s segments (the ASR training chunk
shape), 80 mel, 25 ms / 10 ms.  At N > 1 (torchrun, one rank per GPU) every rank processes its own batch of the same
shape (episode-sharded, weak scaling; the reference's mean is rank-local); value = frames of all
ranks / max time.  The one exchange the path has — the dataset-level statistics all-reduce of configs[4] — is checked
for correctness against the oracle before anything is timed (world > 1) and timed in the `corpus` block.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline       dominant kernel (logmel_ws_kernel) against the SLOWER of the HBM roofline (measured copy bandwidth,
                 960 B/frame) and the FP32 roofline (FMA rate measured in this run, 10 745 flop/frame) — SURVEY.md §8d
  e2e            same metric through the public host-side API (HostPipeline) with pinned HOST buffers, copies inside
                 the timing; primary: int16 PCM in (what the WAV loader tal_asrd_b200.wavio produces, the on-disk
                 format) -> float32 features out; sub-keys: float32 waveforms in, one blocking call, the loader itself
  corpus         configs[4]: 600 one-hour episodes sharded by episode, pass 1 + ONE all-reduce + pass 2 (global CMVN)
  other_configs  configs[0] (one 60 s clip), configs[2] (hour-long episode, one-shot and streamed), configs[3] (ragged)
  cpu_baseline   the reference's own LogMelSpec class (oracle/_ref, built by oracle/build_ref.py) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, SECONDS, SR, HOP, N_MELS = 64, 30, 16000, 160, 80
N_SAMPLES = SECONDS * SR
N_FRAMES = 1 + N_SAMPLES // HOP
FRAMES_PER_STEP = BATCH * N_FRAMES
ALGO_BYTES_PER_FRAME = 4 * HOP + 4 * N_MELS          # each sample read once, each feature written once
ALGO_FLOPS_PER_FRAME = 10745                         # window 400 + rFFT-400 8644 + power 597 + mel 784 + log 160 + mean 160
WORKLOAD = "configs[1]: batch of 64 x 30 s segments, 16 kHz mono, 80 mel, 25 ms/10 ms, batch scalar mean"
FALLBACK_HBM_GBS = 6650.0                            # B200_PROFILING.md fallback
EPISODE_SAMPLES = 57_600_000                         # one hour at 16 kHz
EPISODE_FRAMES = 1 + EPISODE_SAMPLES // HOP


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


# ------------------------------------------------------------------------------------------ the reference on the host cores
def reference_callable():
    """(fn(audio[B, L] float32 torch) -> features, kind).  The reference's own class, verbatim, from oracle/_ref
    (oracle/build_ref.py extracts tal/asr/models.py:15-53 in the build container; torchaudio is in the image), else the
    oracle's fp32 port of the same op sequence (bit-identical to the class in tests/test_oracle.py)."""
    import torch
    try:
        from oracle import build_ref
        cls = build_ref.load()
        if cls is not None:
            mod = cls().eval()
            with torch.no_grad():
                mod(torch.zeros(1, 1600))
            return (lambda x: mod(x)), "reference"
    except Exception:
        pass
    from oracle import logmel_oracle as O
    return (lambda x: O.logmel_port_f32(x)), "port"


def cpu_reference_rate(steps: int, warmup: int, batch_np=None, min_seconds: float = 0.0):
    """Times the reference path on the host cores: `steps` passes over the full 64 x 30 s batch, continued until
    `min_seconds` of timed work have accumulated (at most 300 passes).
    Returns (frames/s, ms/pass, threads, best pass in s, passes, kind)."""
    import torch
    from tal_asrd_b200 import synth
    if batch_np is None:
        batch_np = synth.batch(2020, BATCH, N_SAMPLES)
    x = torch.from_numpy(batch_np)
    if os.environ.get("OMP_NUM_THREADS") in ("1", None) and "LOCAL_RANK" in os.environ:
        torch.set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1; the baseline may use every core
    threads = torch.get_num_threads()
    fn, kind = reference_callable()
    for _ in range(warmup):
        fn(x)
    times = []
    while len(times) < steps or (sum(times) < min_seconds and len(times) < 300):
        t0 = time.perf_counter()
        fn(x)
        times.append(time.perf_counter() - t0)
    mean_s = sum(times) / len(times)
    return FRAMES_PER_STEP / mean_s, mean_s * 1e3, threads, min(times), len(times), kind


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    fps, ms, threads, best, _, kind = cpu_reference_rate(steps, max(1, min(args.warmup, 3)))
    what = ("tal.asr.models.LogMelSpec itself (oracle/_ref/logmelspec.py, extracted verbatim by oracle/build_ref.py)"
            if kind == "reference" else "oracle.logmel_port_f32 (fp32 port of the reference op sequence; oracle/_ref not built)")
    line = {
        "impl": "reference", "metric": "log-mel frames/sec", "value": fps, "unit": "frames/s",
        "realtime_factor": fps * 0.010, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": FRAMES_PER_STEP},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"full 64 x 30 s batch per step, {steps} steps, torch CPU threads={threads}; {what}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ multi-GPU correctness gate
def nccl_corpus_check(mod, dev, world, rank):
    """world > 1, before anything is timed: a five-episode corpus sharded over the ranks, ONE NCCL all-reduce of the
    statistics block, global per-mel CMVN — against the float64 oracle of the whole corpus.  Also the C-ABI
    all-reduce entry point is not needed here: torch.distributed (NCCL) is the transport.  Every rank exits non-zero on
    a mismatch anywhere."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import synth
    from tal_asrd_b200.corpus import corpus_pass, shard_episodes
    lengths = [64000, 48000, 100000, 16001, 80000]
    mine = shard_episodes(lengths, world, rank)
    eps_ = [torch.from_numpy(synth.waveform(1, i, 0, lengths[i])) for i in mine]
    feats, stats = corpus_pass(mod, eps_, norm="row_mel_var", chunk_seconds=2.0)
    torch.cuda.synchronize()
    raw = [O.logmel_unnormalised_f64(synth.waveform(1, i, 0, n)[None])[0] for i, n in enumerate(lengths)]
    allraw = np.concatenate(raw)
    mean, std = allraw.mean(0), allraw.std(0)
    ok = stats.count == allraw.size
    ok = ok and bool(np.allclose(stats.mel_mean.cpu().numpy(), mean, atol=1e-5))
    ok = ok and bool(np.allclose(stats.mel_var.cpu().numpy(), allraw.var(0), rtol=1e-4))
    worst = 0.0
    for f, i in zip(feats, mine):
        want = (raw[i] - mean) / std
        got = f[0].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))))
    ok = ok and worst < 5e-4
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.SUM)
    if int(flag.item()) != 0:
        sys.stderr.write(f"[rank {rank}] NCCL corpus check FAILED (local ok={ok}, worst rel err {worst:.3e})\n")
        dist.destroy_process_group()
        sys.exit(3)
    return {"episodes": len(lengths), "ranks": world, "worst_rel_err_vs_float64_oracle": worst,
            "global_count": stats.count, "status": "pass"}


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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    nccl_check = nccl_corpus_check(mod, dev, world, rank) if world > 1 else None

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
    ms_per_step = max_over_ranks([ev0.elapsed_time(ev1)])[0] / args.steps
    value = world * FRAMES_PER_STEP / (ms_per_step * 1e-3)
    launches_per_step = int(lib.talfe_launches_per_forward(mod.plan(dev).handle, BATCH, N_SAMPLES))

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
    import ctypes
    fma = ctypes.c_double(0.0)
    _lib.check(lib.talfe_probe_fp32_fma_rate(local_rank, ctypes.byref(fma)), "talfe_probe_fp32_fma_rate")
    fp32_tflops = 2.0 * fma.value / 1e12
    kernel_fps = FRAMES_PER_STEP / (kernel_ms * 1e-3)
    hbm_bound_fps = peaks["hbm_gbs"] * 1e9 / ALGO_BYTES_PER_FRAME
    fp32_bound_fps = 2.0 * fma.value / ALGO_FLOPS_PER_FRAME
    achieved_gbs = kernel_fps * ALGO_BYTES_PER_FRAME / 1e9
    achieved_tflops = kernel_fps * ALGO_FLOPS_PER_FRAME / 1e12
    bound = "hbm" if hbm_bound_fps <= fp32_bound_fps else "fp32"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            traffic = json.load(fh).get("logmel_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": bound, "kernel": "logmel_ws_kernel",
        "achieved": achieved_gbs if bound == "hbm" else achieved_tflops,
        "peak": peaks["hbm_gbs"] if bound == "hbm" else fp32_tflops,
        "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
        "frac": kernel_fps / min(hbm_bound_fps, fp32_bound_fps),
        "traffic": traffic, "peak_kind": peak_kind,
        "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                "bound_frames_per_s": hbm_bound_fps, "bytes_per_frame": ALGO_BYTES_PER_FRAME},
        "fp32": {"achieved": achieved_tflops, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_tflops,
                 "bound_frames_per_s": fp32_bound_fps, "flops_per_frame": ALGO_FLOPS_PER_FRAME,
                 "peak_kind": "measured in this run: packed FFMA2 chains on every SM (talfe_probe_fp32_fma_rate), 2 flop per FMA"},
        "step_frac": FRAMES_PER_STEP / (ms_per_step * 1e-3) / min(hbm_bound_fps, fp32_bound_fps),   # whole step (K1 + normalisation), per GPU
        "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step,
        "algorithmic_bytes_per_launch": FRAMES_PER_STEP * ALGO_BYTES_PER_FRAME,
    }

    # ---- end to end through the public API with HOST buffers (tal_asrd_b200.HostPipeline): every step's waveforms go
    # pinned host -> device, through the front end, and its features device -> pinned host, all inside the timed
    # region; consecutive steps overlap on three streams (H2D | transform | D2H), two device slots.  Two distinct
    # host batches alternate.  Primary: int16 PCM, the on-disk format, exactly what tal_asrd_b200.wavio hands over
    # (the reference's loader widens it to float32 on the host, tal/asr/data/util.py:43; here the 1/32768 lives in
    # the kernel's window).  Also: float32 waveforms in, and the same step as ONE blocking call.
    from tal_asrd_b200 import HostPipeline, wavio
    from tal_asrd_b200.hostpipe import bind_host_thread_to_gpu
    numa_cpus = bind_host_thread_to_gpu(local_rank) if world > 1 else None   # one rank per GPU: pin next to its own GPU
    host_in = [torch.empty(BATCH, N_SAMPLES, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        host_in[i].copy_(waves[i].cpu())
    host_pcm = [(h * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory() for h in host_in]
    host_out = [torch.empty(BATCH, N_FRAMES, N_MELS, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = min(50, max(30, args.steps))          # (ramp-up and drain of the pipeline amortise over the run: tools/e2e_depth.py)

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
        dev_ms, wall = max_over_ranks([p0.elapsed_time(p1), wall_ms])
        return dev_ms / e2e_steps, wall / e2e_steps

    pcm_ms, pcm_wall_ms = timed_pipeline(host_pcm)
    f32_ms, _ = timed_pipeline(host_in)

    mod.forward_host(host_in[0], host_out[0], device=dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        mod.forward_host(host_in[i % 2], host_out[i % 2], device=dev)
    e1.record()
    barrier()
    blocking_ms = max_over_ranks([e0.elapsed_time(e1)])[0] / e2e_steps

    loader = None
    if rank == 0:
        # the loader leg of the same path: 64 x 30 s slices out of WAV files (tmpfs) straight into one pinned int16 batch
        try:
            tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            files = []
            pcm_np = host_pcm[0].numpy()
            for f in range(8):
                path = os.path.join(tmp, f"ep{f}.wav")
                wavio.write_wav_pcm16(path, pcm_np[8 * f:8 * f + 8].reshape(-1))     # 8 files of 4 min each
                files.append(path)
            staged = torch.empty(BATCH, N_SAMPLES, dtype=torch.int16).pin_memory()

            def load_batch():
                for b in range(BATCH):
                    wavio.load_audio_segment_pcm16(files[b // 8], (b % 8) * SECONDS, (b % 8 + 1) * SECONDS, out=staged[b])
            load_batch()
            ok = bool(torch.equal(staged, host_pcm[0]))
            t0 = time.perf_counter()
            for _ in range(3):
                load_batch()
            dt = (time.perf_counter() - t0) / 3
            loader = {"api": "wavio.load_audio_segment_pcm16 (reference arguments: path, start_s, end_s) into a pinned int16 batch",
                      "ms_per_batch": dt * 1e3, "frames_per_s_one_host_thread": FRAMES_PER_STEP / dt,
                      "gb_per_s": staged.numel() * 2 / dt / 1e9, "bit_exact_round_trip": ok}
            for p in files:
                os.remove(p)
            os.rmdir(tmp)
        except Exception as exc:                                            # the loader leg is informational
            loader = {"error": repr(exc)}

    e2e = {"value": world * FRAMES_PER_STEP / (pcm_ms * 1e-3), "unit": "frames/s", "ms_per_step": pcm_ms,
           "wall_ms_per_step": pcm_wall_ms,
           "h2d_bytes_per_step": host_pcm[0].numel() * 2, "d2h_bytes_per_step": host_out[0].numel() * 4, "steps": e2e_steps,
           "api": "HostPipeline.submit per step (3 streams, 2 device slots), drain at the end; int16 PCM waveforms in "
                  "(pinned, as tal_asrd_b200.wavio loads them), float32 features out (pinned)",
           "host_cpus_rank0": (f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)} CPUs local to the GPU)" if numa_cpus else "unbound"),
           "f32_input": {"value": world * FRAMES_PER_STEP / (f32_ms * 1e-3), "ms_per_step": f32_ms,
                         "h2d_bytes_per_step": host_in[0].numel() * 4,
                         "api": "the same pipeline fed float32 waveforms (what the reference's loader produces, util.py:43)"},
           "blocking_call": {"value": world * FRAMES_PER_STEP / (blocking_ms * 1e-3), "ms_per_step": blocking_ms,
                             "api": "LogMelSpec.forward_host, float32 in (copy-in, transform, copy-out back to back)"},
           "loader": loader}
    del host_in, host_pcm, host_out

    corpus = None if args.no_extras else corpus_block(mod, lib, _lib, dev, world, rank, barrier, max_over_ranks)
    other = None
    experiments = None
    if rank == 0 and world == 1 and not args.no_extras:
        experiments = kernel_experiments(waves, outs, dev, kernel_ms)
        del waves, outs
        torch.cuda.empty_cache()
        other = other_configs(mod, lib, _lib, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tal_asrd_b200 import synth
        fps, ms, threads, best, passes, kind = cpu_reference_rate(5, 2, synth.batch(2020, BATCH, N_SAMPLES), min_seconds=10.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "ms_per_step": ms,
               "best_ms_per_step": best * 1e3,
               "sample": f"the full 64 x 30 s batch (same samples as the GPU step), mean of {passes} passes "
                         f"(about 10 s of CPU work) after 2 warm-ups; "
                         + ("the reference class itself (oracle/_ref)" if kind == "reference" else "oracle fp32 port")}

    if rank == 0:
        line = {
            "metric": "log-mel frames/sec", "value": value, "unit": "frames/s", "realtime_factor": value * 0.010,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": FRAMES_PER_STEP},     # same dict in both arms
            "config_notes": {"l2_policy": "3 rotating input batches (369 MB) + 2 output buffers (123 MB) > 126 MB L2",
                             "parallelism": f"episode-sharded x{world}, no data-path collective in the step "
                                            f"(the statistics all-reduce of configs[4] is timed in `corpus`)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "corpus": corpus, "other_configs": other,
            "kernel_experiments": experiments, "nccl_check": nccl_check,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def corpus_block(mod, lib, _lib, dev, world, rank, barrier, max_over_ranks):
    """BASELINE configs[4]: 600 one-hour synthetic episodes sharded by episode over the ranks.  Pass 1 transforms every
    episode of this rank and accumulates {count, sum, sumsq, per-mel sums, per-mel sumsq} on the device; the blocks are
    all-reduced ONCE (NCCL over NVLink at world > 1); pass 2 transforms again and applies the GLOBAL per-mel mean /
    variance in place.  A pool of 4 distinct resident episodes per rank (0.9 GB, far larger than L2) stands in for the
    rank's 600 / N episodes, visited in turn: arithmetic and traffic per episode are those of the full corpus."""
    import torch
    from tal_asrd_b200.corpus import CorpusStats, shard_episodes
    EPISODES, POOL, REPS = 600, 4, 2
    L, T = EPISODE_SAMPLES, EPISODE_FRAMES
    mine = shard_episodes([L] * EPISODES, world, rank)
    pool = []
    for i in range(POOL):
        w = torch.empty(1, L, dtype=torch.float32, device=dev)
        _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, 1, L, L, 2020, 10_000 + rank * POOL + i, 0, None))
        pool.append(w)
    out = [torch.empty(1, T, N_MELS, dtype=torch.float32, device=dev) for _ in range(2)]
    blocks = mod.stats_block(dev, rows=len(mine))
    for _ in range(2):
        mod.features(pool[0], norm="row_mel_var", stats=blocks[:1], defer_normalise=True, out=out[0])
    best, total = None, None
    for rep in range(REPS):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for n in range(len(mine)):
            mod.features(pool[n % POOL], norm="row_mel_var", stats=blocks[n:n + 1], defer_normalise=True, out=out[n % 2])
        total = CorpusStats(N_MELS, dev)
        total.add(blocks)
        ev[1].record()
        total.all_reduce()
        ev[2].record()
        for n in range(len(mine)):           # the global statistics are applied inside the transform kernel (no sweep)
            mod.features(pool[n % POOL], norm="row_mel_var", given_stats=total.block, out=out[n % 2])
        ev[3].record()
        barrier()
        t = max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
        if best is None or sum(t) < sum(best):
            best = t
    # The same job when the rank's features stay resident in HBM (600 h = 69 GB of float32 features in total): every
    # episode is transformed ONCE into its own buffer, and after the all-reduce one in-place sweep per episode applies the
    # global statistics — the second transform of the two-pass form above is replaced by 2 x 115 MB of HBM traffic.
    resident = None
    try:
        del out
        torch.cuda.empty_cache()
        need = len(mine) * T * N_MELS * 4
        free, _ = torch.cuda.mem_get_info(dev)
        cannot = max_over_ranks([0.0 if need * 1.15 < free else 1.0])[0]   # every rank takes the same branch (barriers inside)
        if cannot == 0.0:
            feats = torch.empty(len(mine), T, N_MELS, dtype=torch.float32, device=dev)
            feats.zero_()                                                   # pages mapped before the clock starts
            blocks.zero_()
            barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            for n in range(len(mine)):
                mod.features(pool[n % POOL], norm="row_mel_var", stats=blocks[n:n + 1], defer_normalise=True, out=feats[n:n + 1])
            tot2 = CorpusStats(N_MELS, dev)
            tot2.add(blocks)
            ev[1].record()
            tot2.all_reduce()
            ev[2].record()
            for n in range(len(mine)):
                mod.apply_stats(feats[n:n + 1], tot2.block, norm="row_mel_var")
            ev[3].record()
            barrier()
            r1, rr, r2 = max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
            resident = {"feature_bytes_per_rank": need, "transform_and_stats_ms": r1, "allreduce_ms": rr, "apply_sweep_ms": r2,
                        "frames_per_s": float(EPISODES) * T / ((r1 + rr + r2) * 1e-3),
                        "x_realtime": EPISODES * L / SR / ((r1 + rr + r2) * 1e-3),
                        "global_mean": tot2.mean}
            del feats
        else:
            resident = {"skipped": f"needs {need / 1e9:.1f} GB of free device memory, {free / 1e9:.1f} GB available"}
    except Exception as exc:                                                # informational block
        resident = {"error": repr(exc)[:200]}
    del pool
    torch.cuda.empty_cache()
    frames = float(EPISODES) * T
    t1, tr, t2 = best
    hours = EPISODES * L / SR / 3600
    return {"workload": f"configs[4]: {EPISODES} x 1 h synthetic episodes ({hours:.0f} h), sharded by episode over {world} GPU(s), "
                        f"global per-mel CMVN (one statistics all-reduce)",
            "episodes_per_rank_max": -(-EPISODES // world), "frames": frames,
            "pass1_stats_ms": t1, "allreduce_ms": tr, "pass2_normalise_ms": t2,
            "both_passes_frames_per_s": frames / ((t1 + tr + t2) * 1e-3),
            "both_passes_x_realtime": hours * 3600 / ((t1 + tr + t2) * 1e-3),
            "pass1_frames_per_s": frames / (t1 * 1e-3),
            "global_mean": total.mean, "global_count": total.count, "repetitions": REPS, "resident_features": resident,
            "pool": f"{POOL} distinct resident episodes per rank visited in turn (inputs 0.9 GB >> L2)"}


def kernel_experiments(waves, outs, dev, default_kernel_ms):
    """The other formulations of K1 that live in the library, timed on the headline batch in the same run (N = 1):
    `fl` = one thread per frame, tensor memory (tcgen05.st / tcgen05.ld) as the transpose scratch between the FFT stages,
    uniform tables, no inter-warp synchronisation; `legacy` = the homogeneous round-1 kernel.  They are measured
    alternatives, not the shipped path."""
    import torch
    from tal_asrd_b200 import LogMelSpec, frontend
    res = {"default_kernel_ms": default_kernel_ms}
    want = None
    for kern in ("ws", "fl", "legacy"):
        old = os.environ.get("TALFE_KERNEL")
        os.environ["TALFE_KERNEL"] = kern
        saved = dict(frontend._PLANS)
        frontend._PLANS.clear()                                             # plans are cached per table set: a fresh one per kernel
        try:
            m = LogMelSpec().to(dev)
            m.plan(dev)
            for i in range(3):
                m.features(waves[i % len(waves)], norm="none", out=outs[i % 2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(10):
                m.features(waves[i % len(waves)], norm="none", out=outs[i % 2])
            e1.record()
            torch.cuda.synchronize()
            y = m.features(waves[0], norm="none").clone()
            if want is None:
                want = y
            res[kern] = {"kernel_ms": e0.elapsed_time(e1) / 10, "max_abs_diff_vs_ws": float((y - want).abs().max())}
            del m, y
        except Exception as exc:                                            # informational block
            res[kern] = {"error": repr(exc)[:200]}
        finally:
            frontend._PLANS.clear()
            frontend._PLANS.update(saved)
            if old is None:
                os.environ.pop("TALFE_KERNEL", None)
            else:
                os.environ["TALFE_KERNEL"] = old
    return res


def other_configs(mod, lib, _lib, dev):
    """Throughput of the BASELINE configs that are not the headline (N = 1): parity for all of them is in tests/."""
    import numpy as np
    import torch
    from tal_asrd_b200.streaming import DEFAULT_CHUNK_SECONDS, stream_episode

    def fill(rows, n, ep):
        w = torch.empty(rows, n, dtype=torch.float32, device=dev)
        _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, rows, n, n, 2020, ep, 0, None))
        return w

    def timeit(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {}
    # configs[0]: one 60 s clip (the reference's own CPU-runnable case): call latency; also replayed from a CUDA graph
    clip = fill(1, 960000, 500)
    out1 = torch.empty(1, 6001, N_MELS, device=dev)
    ms = timeit(lambda: mod.features(clip, out=out1), 200, warm=10)
    c0 = {"frames": 6001, "us_per_call": ms * 1e3, "frames_per_s": 6001 / (ms * 1e-3),
          "api": "LogMelSpec.features(out=) back to back from Python (host-bound: two launches per call)"}
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            mod.features(clip, out=out1)                                   # workspace for this stream exists before capture
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    mod.features(clip, out=out1)
        torch.cuda.synchronize()
        gms = timeit(g.replay, 20) / 20
        c0["us_per_call_cuda_graph"] = gms * 1e3
        c0["frames_per_s_cuda_graph"] = 6001 / (gms * 1e-3)
    except Exception as exc:
        c0["cuda_graph_error"] = repr(exc)[:200]
    res["configs[0] one 60 s clip"] = c0
    del clip, out1

    # configs[2]: one hour-long episode, one-shot and streamed in overlapping chunks (per-utterance mean over the hour)
    ep = fill(1, EPISODE_SAMPLES, 999)[0]
    T = EPISODE_FRAMES
    c2 = {"frames": T}
    c2["one_shot_ms"] = timeit(lambda: mod(ep[None]), 5)
    c2["streamed_device_resident_ms"] = timeit(lambda: stream_episode(mod, ep), 3)
    c2["streamed_device_resident_30s_chunks_ms"] = timeit(lambda: stream_episode(mod, ep, 30.0), 3)   # (coalesced: nothing to stage on the device)
    c2["streamed_device_resident_30s_chunks_uncoalesced_ms"] = timeit(lambda: stream_episode(mod, ep, 30.0, coalesce_on_device=False), 3)
    eph = ep.cpu().pin_memory()
    c2["streamed_from_pinned_host_ms"] = timeit(lambda: stream_episode(mod, eph, device=dev), 3)
    pcm = (eph * 32768.0).round().to(torch.int16).pin_memory()
    c2["streamed_from_pinned_host_int16_ms"] = timeit(lambda: stream_episode(mod, pcm, device=dev), 3)
    # "per-utterance CMVN" (BASELINE configs[2] wording; the reference itself only removes a scalar mean): per-mel mean and
    # variance over the hour, one extra statistics pass + one in-place sweep behind the transform
    c2["one_shot_cmvn_ms"] = timeit(lambda: mod.features(ep[None], norm="row_mel_var"), 5)
    c2["streamed_device_resident_cmvn_ms"] = timeit(lambda: stream_episode(mod, ep, norm="row_mel_var"), 3)
    c2["streamed_from_pinned_host_int16_cmvn_ms"] = timeit(lambda: stream_episode(mod, pcm, device=dev, norm="row_mel_var"), 3)
    c2["default_chunk_seconds"] = DEFAULT_CHUNK_SECONDS
    c2["one_shot_frames_per_s"] = T / (c2["one_shot_ms"] * 1e-3)
    c2["streamed_from_pinned_host_int16_frames_per_s"] = T / (c2["streamed_from_pinned_host_int16_ms"] * 1e-3)
    res["configs[2] hour-long episode"] = c2
    del ep, eph, pcm
    torch.cuda.empty_cache()

    # configs[3]: ragged batch, 64 utterances log-uniform in 1 s .. 10 min (fixed seed), padding masks
    rng = np.random.default_rng(4)
    lens = np.exp(rng.uniform(np.log(16000), np.log(9_600_000), size=64)).astype(np.int64)
    lens[0], lens[-1] = 16000, 9_600_000
    Lmax = int(lens.max())
    x = torch.zeros(64, Lmax, device=dev)
    for r, n in enumerate(lens):
        _lib.check(lib.talfe_synth_fill(x[r].data_ptr(), _lib.F32, 1, int(n), int(n), 2020, 2000 + r, 0, None))
    lens_t = torch.from_numpy(lens).to(dev)
    valid = int((1 + lens // HOP).sum())
    padded_frames = 64 * (1 + Lmax // HOP)
    c3 = {"rows": 64, "valid_frames": valid, "padded_frames": padded_frames}
    c3["padded_reference_semantics_ms"] = timeit(lambda: mod(x), 5)
    # the same result (reference semantics: every row padded with zeros, padding frames in the mean) with the collater's
    # lengths as a hint: only frames that can see a real sample are computed, the rest is written once as log(eps) - mean
    c3["padded_reference_semantics_with_lens_hint_ms"] = timeit(lambda: mod.features(x, audio_lens=lens_t, lens_are_padding=True), 5)
    c3["lens_hint_max_abs_diff"] = float((mod.features(x, audio_lens=lens_t, lens_are_padding=True) - mod(x)).abs().max())
    c3["per_row_ms"] = timeit(lambda: mod.features(x, audio_lens=lens_t, norm="row"), 5)
    c3["packed_ms"] = timeit(lambda: mod.features_packed(x, lens_t, norm="row"), 5)
    c3["padded_frames_per_s"] = padded_frames / (c3["padded_reference_semantics_ms"] * 1e-3)
    c3["packed_valid_frames_per_s"] = valid / (c3["packed_ms"] * 1e-3)
    res["configs[3] ragged batch 1 s .. 10 min"] = c3
    del x, lens_t
    torch.cuda.empty_cache()

    # the rest of the constructor's signature: LogMelSpec(sr != 16000) runs the generic-geometry kernel (a plain direct-DFT
    # formulation, correct for any n_fft / hop; the reference itself only ever runs at 16 kHz)
    from tal_asrd_b200 import LogMelSpec
    rates = {}
    for sr in (8000, 22050, 48000):
        m = LogMelSpec(sr=sr).to(dev)
        xs = torch.randn(16, 10 * sr, device=dev) * 0.1
        ys = m(xs)
        ms = timeit(lambda: m.features(xs, out=ys), 5)
        rates[str(sr)] = {"n_fft": m.n_fft, "hop": m.hop, "batch": "16 x 10 s", "frames": int(ys.shape[0] * ys.shape[1]),
                          "ms": ms, "frames_per_s": ys.shape[0] * ys.shape[1] / (ms * 1e-3)}
        del m, xs, ys
    res["other sample rates (generic kernel)"] = rates
    return res


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
    ap.add_argument("--no-extras", action="store_true", help="skip the corpus / other_configs blocks (profiling runs)")
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
