"""Hour-long episodes streamed through the front end in overlapping chunks (BASELINE config 3).

Reference behaviour being matched: ``tal/baseline/reconcile.py:76-85`` pushes a WHOLE episode
(~57.6 M samples) through ``model.encode`` -> ``LogMelSpec.forward`` as one ``[1, L]`` row, so the
scalar mean is over the entire episode and reflection happens only at the true ends.  Streaming must
therefore produce exactly the one-shot frame grid: chunk boundaries at multiples of the hop, a
200-sample halo on both sides, reflection only at sample 0 and sample L-1, and the mean applied
after the last chunk (SURVEY.md §5 "long-context", §7 "streaming equivalence").
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .frontend import HOP, N_FFT, LogMelSpec, _DTYPES, _NORMS, num_frames

HALO = N_FFT // 2


def chunk_plan(total_len: int, chunk_frames: int):
    """[(frame0, frame1, sample_lo, sample_hi)]: frames [frame0, frame1) need samples [lo, hi)."""
    T = num_frames(total_len)
    plan = []
    for f0 in range(0, T, chunk_frames):
        f1 = min(T, f0 + chunk_frames)
        lo, hi = HOP * f0 - HALO, HOP * (f1 - 1) + HALO      # samples the frames span before reflection
        need_lo, need_hi = lo, hi
        if hi > total_len:                                    # right edge reflects back to 2(L-1) - g
            need_lo = min(need_lo, 2 * (total_len - 1) - (hi - 1))
        if lo < 0:                                            # left edge reflects forward to -g
            need_hi = max(need_hi, -lo + 1)
        plan.append((f0, f1, max(0, need_lo), min(total_len, need_hi)))
    return plan


DEFAULT_CHUNK_SECONDS = 300.0      # 30 001 frames = 938 tiles: every persistent CTA gets >= 6 tiles per chunk, and the per-chunk
                                   # launch cost (two launches, ~10 us of device latency) stays below 10 % of the chunk's work


MIN_DEVICE_CHUNK_FRAMES = 148 * 16 * 32   # a device-resident episode: at least 16 tiles per persistent CTA and launch (~758 s)


def stream_episode(frontend: LogMelSpec, episode: torch.Tensor, chunk_seconds: float = DEFAULT_CHUNK_SECONDS,
                   device: Optional[torch.device] = None, norm: str = "batch",
                   out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
                   normalise: bool = True, coalesce_on_device: bool = True) -> torch.Tensor:
    """episode: 1-D waveform (float32 / float16 / int16), on the host (pinned for full copy speed) or
    already on the device.  Returns [1, T, n_mels] float32 on the device, equal to
    ``frontend(episode[None])`` for norm='batch'.

    The chunk loop runs natively (``talfe_stream_episode``): host->device copies of chunk k+1 go out on a side
    stream while chunk k is being transformed (two staging buffers; a device-resident episode is transformed
    where it lies), statistics accumulate on the device across chunks and one in-place sweep at the end applies
    them.  With ``normalise=False`` the features are left un-normalised and ``stats`` holds the sums (for a
    dataset-level all-reduce, see corpus.py).

    ``chunk_seconds`` sizes the HOST staging; an episode that already lies on the device has nothing to stage, so its
    chunks are coalesced to at least ``MIN_DEVICE_CHUNK_FRAMES`` frames per launch (small chunks only add launches:
    120 chunks of 30 s cost 1.9 ms against 0.18 ms for the hour in one launch); ``coalesce_on_device=False`` keeps the
    requested chunk size (tests of the chunk-boundary logic).  The un-normalised features do not depend on the chunking
    (every frame depends on its own 400 samples only).

    Buffer lifetime: when the call returns every copy out of a HOST ``episode`` has completed (the transforms may
    still be running on the current stream), so the caller may reuse or drop the host tensor immediately.
    """
    if episode.dim() != 1:
        raise ValueError("episode must be a 1-D waveform")
    if (frontend.n_fft, frontend.hop) != (N_FFT, HOP):
        raise NotImplementedError("chunked streaming exists for the reference's 16 kHz geometry only (n_fft 400, hop 160); "
                                  "call the module on the whole episode for other sample rates")
    if episode.dtype not in _DTYPES:
        episode = episode.float()
    if not episode.is_contiguous():
        episode = episode.contiguous()
    L = episode.numel()
    T = num_frames(L)
    if device is None:
        device = episode.device if episode.is_cuda else torch.device("cuda", torch.cuda.current_device())
    elif episode.is_cuda and episode.device != torch.device(device):
        raise ValueError("a device-resident episode must live on the device it is transformed on")
    plan = frontend.plan(device)
    M = frontend.n_mels
    chunk_frames = max(1, int(round(chunk_seconds * frontend.sr / HOP)))
    if episode.is_cuda and coalesce_on_device:
        chunk_frames = max(chunk_frames, MIN_DEVICE_CHUNK_FRAMES)
    with torch.no_grad(), torch.cuda.device(device):
        if out is None:
            out = torch.empty(1, T, M, dtype=torch.float32, device=device)
        if stats is None:
            stats = frontend.stats_block(device)
        compute = torch.cuda.current_stream(device)
        lib = plan.lib
        code = _DTYPES[episode.dtype]
        staging = None
        if not episode.is_cuda:
            staging = torch.empty(int(lib.talfe_stream_staging_bytes(code, chunk_frames)), dtype=torch.uint8, device=device)
        workspace = plan.workspace(compute.cuda_stream, 1, min(chunk_frames, T))
        _lib.check(lib.talfe_stream_episode(plan.handle, episode.data_ptr(), code, L, chunk_frames, out.data_ptr(),
                                            _NORMS[norm], 0 if normalise else 1, stats.data_ptr(), frontend.eps,
                                            staging.data_ptr() if staging is not None else None,
                                            staging.numel() if staging is not None else 0,
                                            workspace.data_ptr(), workspace.numel(),
                                            compute.cuda_stream), "talfe_stream_episode")
        # every side-stream copy into `staging` is followed (through an event) by a transform already queued on
        # `compute`, so handing the block back to the caching allocator here is ordered correctly
    return out
