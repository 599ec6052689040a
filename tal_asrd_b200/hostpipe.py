"""Batches that live in HOST memory pushed through the front end with the PCIe copies of consecutive
batches overlapped (pinned staging -> async H2D double-buffering).

The reference moves every batch to the GPU with a blocking ``.cuda()`` / Lightning's transfer hook and
reads results back with ``.cpu()`` (``tal/asr/transcribe.py:179-191``, ``tal/baseline/reconcile.py:76-85``):
copy-in, compute and copy-out of one batch run back to back and the link idles in one direction at a
time.  ``HostPipeline`` keeps ``depth`` device-side slots and three streams (H2D | front end | D2H), so
that while batch n is being transformed and batch n-1 is on its way back, batch n+1 is already coming
in; PCIe is full duplex, so a steady stream of batches costs max(H2D, D2H) per batch instead of their
sum.  Each batch is still exactly ``LogMelSpec.forward`` on that batch (its own scalar mean).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch

from .frontend import LogMelSpec, num_frames


def bind_host_thread_to_gpu(device_index: int) -> Optional[List[int]]:
    """Restricts the calling process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    pinned buffers allocated afterwards land in memory next to that GPU's PCIe root.  On multi-socket hosts with one
    rank per GPU this keeps H2D/D2H traffic off the inter-socket link.  Returns the CPU list, or None when NVML or
    the affinity call is unavailable (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


class HostPipeline:
    """pipe = HostPipeline(frontend, device); ev = pipe.submit(audio_host, out_host); ...; pipe.drain()

    ``audio_host``  [B, L] float32 / float16 / int16 on the host (pinned for full copy speed; pageable works
                    but the copy then blocks the calling thread)
    ``out_host``    [B, 1 + L // 160, n_mels] float32 on the host (pinned), filled when the returned event
                    (or ``drain()``) has completed
    Calls are asynchronous with respect to the host: ``submit`` returns as soon as the work is queued.
    The caller must not touch ``audio_host`` / ``out_host`` of a batch before its event has completed.
    """

    def __init__(self, frontend: LogMelSpec, device: Optional[torch.device] = None, depth: int = 2,
                 norm: str = "batch", layout: str = "tm"):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device: the front end has no CPU implementation")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.frontend, self.depth, self.norm, self.layout = frontend, depth, norm, layout
        self.h2d = torch.cuda.Stream(self.device)
        self.compute = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._x = [None] * depth          # device waveforms per slot
        self._y = [None] * depth          # device features per slot
        self._ev_in = [None] * depth      # H2D of the slot's batch finished
        self._ev_done = [None] * depth    # the front end has finished reading x / writing y of the slot
        self._ev_out = [None] * depth     # D2H of the slot's batch finished
        self._n = 0
        frontend.plan(self.device)        # tables uploaded before the first batch is timed

    def _slot_buffers(self, s: int, audio_host: torch.Tensor, out_shape) -> Tuple[torch.Tensor, torch.Tensor]:
        x, y = self._x[s], self._y[s]
        if x is None or x.shape != audio_host.shape or x.dtype != audio_host.dtype:
            # a replaced buffer may still be in flight on a side stream: the caching allocator must not hand it out
            # again before those streams are past it
            for old in (x, y):
                if old is not None:
                    for st in (self.h2d, self.compute, self.d2h):
                        old.record_stream(st)
            x = torch.empty(audio_host.shape, dtype=audio_host.dtype, device=self.device)
            y = torch.empty(out_shape, dtype=torch.float32, device=self.device)
            self._x[s], self._y[s] = x, y
        return x, y

    def submit(self, audio_host: torch.Tensor, out_host: torch.Tensor,
               audio_lens: Optional[torch.Tensor] = None) -> torch.cuda.Event:
        if audio_host.is_cuda or out_host.is_cuda:
            raise ValueError("HostPipeline moves HOST batches; call LogMelSpec.forward for device tensors")
        if audio_host.dim() != 2:
            raise ValueError(f"audio must be [batch, audio_len], got shape {tuple(audio_host.shape)}")
        B, L = audio_host.shape
        T, M = num_frames(L, self.frontend.n_fft, self.frontend.hop), self.frontend.n_mels
        shape = (B, T, M) if self.layout == "tm" else (B, M, T)
        if tuple(out_host.shape) != shape or out_host.dtype != torch.float32 or not out_host.is_contiguous():
            raise ValueError(f"out_host must be a contiguous float32 tensor of shape {shape}")
        s = self._n % self.depth
        self._n += 1
        with torch.no_grad(), torch.cuda.device(self.device):
            x, y = self._slot_buffers(s, audio_host, shape)
            with torch.cuda.stream(self.h2d):
                if self._ev_done[s] is not None:
                    self.h2d.wait_event(self._ev_done[s])        # the previous batch of this slot has been transformed
                x.copy_(audio_host, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ev_in)
                if self._ev_out[s] is not None:
                    self.compute.wait_event(self._ev_out[s])     # y of this slot has left for the host
                self.frontend.features(x, audio_lens=audio_lens, norm=self.norm, layout=self.layout, out=y)
                ev_done = torch.cuda.Event()
                ev_done.record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(ev_done)
                out_host.copy_(y, non_blocking=True)
                ev_out = torch.cuda.Event()
                ev_out.record(self.d2h)
            self._ev_in[s], self._ev_done[s], self._ev_out[s] = ev_in, ev_done, ev_out
        return ev_out

    def drain(self) -> None:
        """Blocks until every submitted batch has landed in its ``out_host``."""
        self.d2h.synchronize()
        self.compute.synchronize()
        self.h2d.synchronize()

    def run(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> int:
        """Submits every (audio_host, out_host) pair and drains; returns the number of batches."""
        n = 0
        for audio_host, out_host in batches:
            self.submit(audio_host, out_host)
            n += 1
        self.drain()
        return n
