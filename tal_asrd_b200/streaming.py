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
from .frontend import HOP, N_FFT, LogMelSpec, _DTYPES, _NORMS, _require_cuda, _run, num_frames

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


def stream_episode(frontend: LogMelSpec, episode: torch.Tensor, chunk_seconds: float = 30.0,
                   device: Optional[torch.device] = None, norm: str = "batch",
                   out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
                   normalise: bool = True) -> torch.Tensor:
    """episode: 1-D waveform (float32 / float16 / int16), on the host (pinned for full copy speed) or
    already on the device.  Returns [1, T, n_mels] float32 on the device, equal to
    ``frontend(episode[None])`` for norm='batch'.

    Host->device copies of chunk k+1 run on a side stream while chunk k is being transformed
    (two staging buffers); statistics accumulate on the device across chunks; one in-place sweep at
    the end applies them.  With ``normalise=False`` the features are left un-normalised and ``stats``
    holds the sums (for a dataset-level all-reduce, see corpus.py).
    """
    if episode.dim() != 1:
        raise ValueError("episode must be a 1-D waveform")
    if episode.dtype not in _DTYPES:
        episode = episode.float()
    L = episode.numel()
    T = num_frames(L)
    if device is None:
        device = episode.device if episode.is_cuda else torch.device("cuda", torch.cuda.current_device())
    plan = frontend.plan(device)
    M = frontend.n_mels
    chunk_frames = max(1, int(round(chunk_seconds * frontend.sr / HOP)))
    with torch.no_grad(), torch.cuda.device(device):
        if out is None:
            out = torch.empty(1, T, M, dtype=torch.float32, device=device)
        if stats is None:
            stats = frontend.stats_block(device)
        chunks = chunk_plan(L, chunk_frames)
        compute = torch.cuda.current_stream(device)
        on_device = episode.is_cuda
        if not on_device and episode.is_contiguous():
            # native loop (talfe_stream_episode): copies and transforms are enqueued from C, chunk after chunk
            lib = plan.lib
            code = _DTYPES[episode.dtype]
            staging = torch.empty(int(lib.talfe_stream_staging_bytes(code, chunk_frames)), dtype=torch.uint8, device=device)
            ws_bytes = plan.workspace_bytes(1, min(chunk_frames, T))
            workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            _lib.check(lib.talfe_stream_episode(plan.handle, episode.data_ptr(), code, L, chunk_frames, out.data_ptr(),
                                                _NORMS[norm], 0 if normalise else 1, stats.data_ptr(), frontend.eps,
                                                staging.data_ptr(), staging.numel(), workspace.data_ptr(), ws_bytes,
                                                compute.cuda_stream), "talfe_stream_episode")
            # every side-stream copy into `staging` is followed (through an event) by a transform already queued on
            # `compute`, so handing the blocks back to the caching allocator here is ordered correctly
            return out
        if not on_device:
            copy_stream = torch.cuda.Stream(device)
            max_n = max(hi - lo for _, _, lo, hi in chunks)
            staging = [torch.empty(max_n, dtype=episode.dtype, device=device) for _ in range(2)]
            free = [None, None]
        # 'batch' over a single row and 'row' coincide; per-mel modes need the column sums too
        run_norm = _NORMS[norm] if norm != "none" else _lib.NORM_NONE
        for k, (f0, f1, lo, hi) in enumerate(chunks):
            n = hi - lo
            if on_device:
                buf = episode[lo:hi]
            else:
                buf = staging[k % 2][:n]
                with torch.cuda.stream(copy_stream):
                    if free[k % 2] is not None:
                        copy_stream.wait_event(free[k % 2])
                    buf.copy_(episode[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(copy_stream)
                compute.wait_event(ready)
            _run(plan, buf.unsqueeze(0), norm=run_norm, layout=_lib.LAYOUT_TM, eps=frontend.eps, lens=None,
                 origin=lo, total_len=L, frame0=f0, n_frames=f1 - f0, out=out[:, f0:f1],
                 stats=stats, accumulate=(k > 0), defer=True)
            if not on_device:
                free[k % 2] = torch.cuda.Event()
                free[k % 2].record(compute)
        if normalise and norm != "none":
            frontend.apply_stats(out, stats, norm=norm)
    return out
