"""Host-side mirror of the reference's feature-extraction interface.

``LogMelSpec`` here has the constructor, call signature, output shape/dtype/contiguity, error
behaviour and (for checkpoint compatibility) buffer names of
``/root/reference/tal/asr/models.py:15-53``; swap it in with

    import tal.asr.models as ref_models
    ref_models.LogMelSpec = tal_asrd_b200.LogMelSpec          # before constructing ASRModel / SDModel
    # or, on an existing model:   model.logmelspec = tal_asrd_b200.LogMelSpec(n_mels=80)

and ``ASRModel.extract_features`` (models.py:154-162) / ``SDModel.extract_features`` (:430-438)
run unchanged.  All arithmetic happens in the sm_100a kernels of ``csrc/talfe.cu`` behind the C ABI
in ``include/talfe.h``; PyTorch only provides device memory and the current stream.
"""
from __future__ import annotations

import hashlib
import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

DEFAULT_SR = 16000          # /root/reference/tal/asr/data/__init__.py:6
HOP = 160
N_FFT = 400
_NORMS = {"none": _lib.NORM_NONE, "batch": _lib.NORM_BATCH_MEAN, "row": _lib.NORM_ROW_MEAN,
          "row_mel": _lib.NORM_ROW_MEL_MEAN, "row_mel_var": _lib.NORM_ROW_MEL_MEANVAR}
_LAYOUTS = {"tm": _lib.LAYOUT_TM, "mt": _lib.LAYOUT_MT}
_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.int16: _lib.I16}


def num_frames(n_samples: int, n_fft: int = N_FFT, hop: int = HOP) -> int:
    """Frames torch.stft(center=True, reflect) makes: 1 + (L + 2 (n_fft // 2) - n_fft) // hop — 1 + L // 160 for the
    reference's geometry; RuntimeError for L <= n_fft // 2 like the reference's reflect pad."""
    if n_samples <= n_fft // 2:
        raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, "
                           f"but got: padding ({n_fft // 2}, {n_fft // 2}) at dimension 2 of input of length {n_samples}")
    return 1 + (n_samples + 2 * (n_fft // 2) - n_fft) // hop


def geometry(sr: int):
    """(n_fft, hop) the reference derives from its sample rate (tal/asr/models.py:24-32)."""
    return int(25 / 1000 * sr), int(10 / 1000 * sr)


def reference_tables(n_mels: int = 80, sr: int = DEFAULT_SR):
    """(window[n_fft], fb[n_fft // 2 + 1, n_mels]) evaluated with the same fp32 torch ops as the reference's
    buffers (torch.hann_window; torchaudio's HTK filterbank recipe with f_max = sr // 2), hence bit-identical to them."""
    n_fft = int(25 / 1000 * sr)
    window = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    freqs = torch.linspace(0, sr // 2, n_fft // 2 + 1)
    top = 2595.0 * math.log10(1.0 + (sr // 2) / 700.0)
    mel_pts = torch.linspace(0.0, top, n_mels + 2)
    hz_pts = 700.0 * (10.0 ** (mel_pts / 2595.0) - 1.0)
    width = hz_pts[1:] - hz_pts[:-1]
    delta = hz_pts.unsqueeze(0) - freqs.unsqueeze(1)
    fb = torch.clamp(torch.minimum((-1.0 * delta[:, :-2]) / width[:-1], delta[:, 2:] / width[1:]), min=0.0)
    return window, fb.contiguous()


def encoder_padding_mask(audio_lens: torch.Tensor, enc_frames: int) -> torch.Tensor:
    """Vectorised form of the mask loop in ``ASRModel.encode_features`` (tal/asr/models.py:178-187):
    True (= ignore) where the encoder frame index is >= audio_len // (audio_lens.max() // enc_frames).
    Stays on the device of ``audio_lens``; no per-row host loop."""
    scaled = audio_lens // (audio_lens.max() // enc_frames)
    return torch.arange(enc_frames, device=audio_lens.device)[None, :] >= scaled[:, None]


class _Plan:
    """Owns one talfe_plan (device tables) and frees it with the object.  Also keeps the per-stream workspaces the
    calls need (allocated once and grown on demand, not once per call)."""

    def __init__(self, device: torch.device, n_mels: int, window: torch.Tensor, fb: torch.Tensor, hop: int = HOP):
        import ctypes
        self.lib = _lib.load()
        self.handle = ctypes.c_void_p()
        self.n_mels = n_mels
        self.device = device
        win = window.detach().to("cpu", torch.float32).contiguous()
        fbc = fb.detach().to("cpu", torch.float32).contiguous()
        self.n_fft, self.hop = int(win.numel()), int(hop)
        if tuple(fbc.shape) != (self.n_fft // 2 + 1, n_mels):
            raise ValueError("fb must be [n_fft // 2 + 1, n_mels] for a window of n_fft elements")
        # n_fft 400 / hop 160 -> the specialised kernels; any other geometry -> the generic kernel (talfe_generic.cuh)
        if (self.n_fft, self.hop) == (N_FFT, HOP):
            _lib.check(self.lib.talfe_plan_create(ctypes.byref(self.handle), device.index, n_mels,
                                                  win.data_ptr(), fbc.data_ptr()), "talfe_plan_create")
        else:
            _lib.check(self.lib.talfe_plan_create_ex(ctypes.byref(self.handle), device.index, self.n_fft, self.hop, n_mels,
                                                     win.data_ptr(), fbc.data_ptr()), "talfe_plan_create_ex")
        self._ws = {}                 # cuda_stream -> uint8 workspace tensor
        self._ws_need = {}            # (batch, n_frames) -> bytes
        self._run = self.lib.talfe_run
        self._forward = self.lib.talfe_logmel_forward

    def workspace_bytes(self, batch: int, n_frames: int) -> int:
        key = (batch, n_frames)
        need = self._ws_need.get(key)
        if need is None:
            need = self._ws_need[key] = int(self.lib.talfe_workspace_bytes(self.handle, batch, n_frames))
        return need

    def workspace(self, stream_ptr: int, batch: int, n_frames: int) -> torch.Tensor:
        """Workspace for a call on the given stream.  One buffer per stream (calls on one stream are ordered, so they
        can share it); it only ever grows.  The replaced buffer goes back to the caching allocator, which keeps it
        away from other streams until the work queued on this one has passed."""
        need = self.workspace_bytes(batch, n_frames)
        ws = self._ws.get(stream_ptr)
        if ws is None or ws.numel() < need:
            ws = self._ws[stream_ptr] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=self.device)
        return ws

    def __del__(self):
        try:
            if self.handle:
                self.lib.talfe_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# Plans hold native handles (a ctypes CDLL and a talfe_plan*): they live HERE, outside every module's state, so that a
# LogMelSpec can be deep-copied, pickled and torch.save'd like the reference module at any time, and so that copies of
# a module (EMA / SWA replicas, DDP spawn) share one set of device tables.  Key: device, n_mels and the table values.
_PLANS = {}


def _shared_plan(device: torch.device, n_mels: int, window: torch.Tensor, fb: torch.Tensor, hop: int = HOP) -> _Plan:
    win = window.detach().to("cpu", torch.float32).contiguous()
    fbc = fb.detach().to("cpu", torch.float32).contiguous()
    # the library reads its development switches when a plan is created: they are part of what a plan is
    knobs = tuple(os.environ.get(k) for k in ("TALFE_KERNEL", "TALFE_L2_PREFETCH", "TALFE_FUSED_NORM", "TALFE_TMA", "TALFE_LIB"))
    key = (device.index, n_mels, hop, hashlib.sha1(win.numpy().tobytes()).digest(), hashlib.sha1(fbc.numpy().tobytes()).digest(), knobs)
    plan = _PLANS.get(key)
    if plan is None:
        plan = _PLANS[key] = _Plan(device, n_mels, win, fbc, hop)
    return plan


def _require_cuda(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise RuntimeError("tal_asrd_b200 front end runs only on a CUDA device (sm_100a); "
                           "there is no CPU implementation — move the waveform to the GPU")
    return t.device


def _run(plan: _Plan, audio: torch.Tensor, *, norm: int, layout: int, eps: float, lens: Optional[torch.Tensor],
         origin: int = 0, total_len: Optional[int] = None, frame0: int = 0, n_frames: Optional[int] = None,
         out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None, accumulate: bool = False,
         defer: bool = False, out_offsets: Optional[torch.Tensor] = None, packed_frames: int = 0,
         bands=None, given_stats: Optional[torch.Tensor] = None, padding_hint: bool = False) -> torch.Tensor:
    device = audio.device
    if torch.cuda.current_device() != device.index:
        with torch.cuda.device(device):          # launches and allocations below need `device` current
            return _run(plan, audio, norm=norm, layout=layout, eps=eps, lens=lens, origin=origin, total_len=total_len,
                        frame0=frame0, n_frames=n_frames, out=out, stats=stats, accumulate=accumulate, defer=defer,
                        out_offsets=out_offsets, packed_frames=packed_frames, bands=bands, given_stats=given_stats, padding_hint=padding_hint)
    B, buf_len = audio.shape
    if total_len is None:
        total_len = buf_len
    if n_frames is None:
        n_frames = num_frames(total_len, plan.n_fft, plan.hop) - frame0
    M = plan.n_mels
    shape = (B, n_frames, M) if layout == _lib.LAYOUT_TM else (B, M, n_frames)
    if out_offsets is not None:
        shape = (packed_frames, M)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
        raise ValueError(f"out must be a contiguous float32 tensor of shape {shape} on {device}")
    stream_ptr = torch.cuda.current_stream(device).cuda_stream
    workspace = plan.workspace(stream_ptr, B, n_frames)
    job = _lib.Job()
    job.wave = audio.data_ptr()
    job.wave_dtype = _DTYPES[audio.dtype]
    job.norm = norm
    job.batch = B
    job.row_stride = audio.stride(0) if B > 1 else max(audio.stride(0), buf_len)
    job.buf_len = buf_len
    job.origin = origin
    job.total_len = total_len
    job.lens = lens.data_ptr() if lens is not None else None
    job.frame0 = frame0
    job.n_frames = n_frames
    job.out = out.data_ptr()
    job.out_row_stride = 0
    job.out_layout = layout
    job.accumulate_stats = 1 if accumulate else 0
    job.eps = eps
    job.defer_normalise = 1 if defer else 0
    job.stats = stats.data_ptr() if stats is not None else None
    job.workspace = workspace.data_ptr()
    job.workspace_bytes = workspace.numel()
    job.out_offsets = out_offsets.data_ptr() if out_offsets is not None else None
    if bands is not None:
        fb, tb = bands
        job.freq_bands, job.time_bands, job.n_bands = fb.data_ptr(), tb.data_ptr(), fb.shape[1]
    if given_stats is not None:
        job.given_stats = given_stats.data_ptr()
    job.lens_are_padding_hint = 1 if padding_hint else 0
    job.stream = stream_ptr
    rc = plan._run(plan.handle, job)
    if rc:
        _lib.check(rc, "talfe_run")
    return out


def _forward_fast(plan: _Plan, audio: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LogMelSpec.forward semantics through the one-call entry point talfe_logmel_forward (no job struct to fill):
    the shortest host path, which is what bounds small calls (one 60 s clip is ~10 us of device work)."""
    device = audio.device
    if torch.cuda.current_device() != device.index:
        with torch.cuda.device(device):
            return _forward_fast(plan, audio, eps, out)
    B, L = audio.shape
    T = num_frames(L, plan.n_fft, plan.hop)
    if out is None:
        out = torch.empty((B, T, plan.n_mels), dtype=torch.float32, device=device)
    stream_ptr = torch.cuda.current_stream(device).cuda_stream
    ws = plan.workspace(stream_ptr, B, T)
    rc = plan._forward(plan.handle, audio.data_ptr(), _DTYPES[audio.dtype], B, L,
                       audio.stride(0) if B > 1 else max(audio.stride(0), L), out.data_ptr(), eps,
                       ws.data_ptr(), ws.numel(), stream_ptr)
    if rc:
        _lib.check(rc, "talfe_logmel_forward")
    return out


def _prepare_audio(audio: torch.Tensor) -> torch.Tensor:
    if audio.dim() != 2:
        raise ValueError(f"audio must be [batch, audio_len] (tal/asr/models.py:36-40), got shape {tuple(audio.shape)}")
    if audio.dtype not in _DTYPES:
        audio = audio.float()
    if audio.stride(-1) != 1 or (audio.shape[0] > 1 and audio.stride(0) < audio.shape[1]):
        audio = audio.contiguous()
    return audio


class _Buffers(nn.Module):
    def __init__(self, name: str, value: torch.Tensor):
        super().__init__()
        self.register_buffer(name, value)


class _MelTransformBuffers(nn.Module):
    """Keeps the reference's state_dict keys alive:
    ``mel_transform.spectrogram.window`` and ``mel_transform.mel_scale.fb``."""

    def __init__(self, window: torch.Tensor, fb: torch.Tensor):
        super().__init__()
        self.spectrogram = _Buffers("window", window)
        self.mel_scale = _Buffers("fb", fb)


class LogMelSpec(nn.Module):
    """Drop-in for ``tal.asr.models.LogMelSpec`` (models.py:15-53) on a B200.

    forward(audio[B, L]) -> [B, 1 + L // 160, n_mels], contiguous, no grad:
    log(mel_power + eps) minus ONE scalar mean over the whole batch result (models.py:50-52).
    Input may be float32, float16 (what ``.half()`` callers pass, system.py:92) or int16 PCM.

    Output dtype follows the reference's type promotion between the waveform and the module's buffers
    (``window`` / ``fb``): float32 whenever either is float32 — a float32 module fed ``audio.half()`` returns
    float32, as torchaudio does — and float16 only when the module itself has been halved (``model.half()``,
    tal/asr/transcribe.py:255) AND the waveform is half, which is what the half-precision TDS encoder then
    expects.  The arithmetic is float32 in every case, with the unrounded float32 tables (a halved module's
    rounded buffers are kept for ``state_dict`` compatibility only); the cast happens once at the end.
    """

    def __init__(self, sr: int = DEFAULT_SR, n_mels: int = 80, eps: float = 1e-6, detect_padding: bool = False):
        super().__init__()
        # detect_padding (extension, off by default = the reference's constructor): forward() first finds each row's last
        # non-zero sample on the device (talfe_detect_padding: the collaters zero-pad, tal/asr/data/aligned.py:246-270) and
        # then computes only the frames that can see a real sample — the same result, faster on heavily padded batches,
        # one small extra launch on dense ones
        self.detect_padding = bool(detect_padding)
        # sr = 16000 (n_fft 400, hop 160: the only rate the reference runs at, tal/asr/data/__init__.py:6) is served by the
        # specialised kernels; any other rate by the generic kernel (same C ABI, same semantics, csrc/talfe_generic.cuh)
        self.n_fft, self.hop = geometry(sr)
        if self.n_fft < 2 or self.hop < 1 or self.n_fft > 1280:
            raise NotImplementedError(f"sr={sr}: n_fft={self.n_fft} is outside what the sm_100a kernels stage in shared memory (2..1280)")
        if not 1 <= n_mels <= 80:
            raise NotImplementedError("n_mels must be in 1..80")
        window, fb = reference_tables(n_mels, sr)
        self.mel_transform = _MelTransformBuffers(window, fb)
        self.sr, self.n_mels, self.eps = sr, n_mels, eps
        self._plans = {}

    # --- native handles never enter the module's state: deepcopy / pickle / torch.save work at any time (the
    # reference module can be copied and pickled; Lightning's ddp spawn and EMA / SWA replicas rely on it)
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_plans"] = {}
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._plans = {}

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._plans = {}                       # tables may have been replaced by checkpoint values

    def _apply(self, fn, *args, **kwargs):
        self._plans = {}
        return super()._apply(fn, *args, **kwargs)

    def plan(self, device: torch.device) -> _Plan:
        key = device.index
        plan = self._plans.get(key)
        if plan is None:
            window, fb = self.mel_transform.spectrogram.window, self.mel_transform.mel_scale.fb
            if window.dtype != torch.float32 or fb.dtype != torch.float32:
                # halved module: its buffers hold ROUNDED tables; the kernels run in float32 from the exact ones
                window, fb = reference_tables(self.n_mels, self.sr)
            plan = self._plans[key] = _shared_plan(device, self.n_mels, window, fb, self.hop)
        return plan

    def _out_dtype(self, audio: torch.Tensor) -> torch.dtype:
        if audio.dtype == torch.float16 and self.mel_transform.spectrogram.window.dtype == torch.float16:
            return torch.float16
        return torch.float32

    @torch.jit.ignore
    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            audio = _prepare_audio(audio)
            device = _require_cuda(audio)
            num_frames(audio.shape[1], self.n_fft, self.hop)
            if self.detect_padding:
                y = self.features(audio, audio_lens=self.padding_lens(audio), lens_are_padding=True)
            else:
                y = _forward_fast(self.plan(device), audio, self.eps)
            return y if self._out_dtype(audio) == torch.float32 else y.half()

    @torch.jit.ignore
    def padding_lens(self, audio: torch.Tensor) -> torch.Tensor:
        """int64 [B] on the device: 1 + index of every row's last non-zero sample (0 for an all-zero row)."""
        audio = _prepare_audio(audio)
        device = _require_cuda(audio)
        plan = self.plan(device)
        lens = torch.empty(audio.shape[0], dtype=torch.int64, device=device)
        with torch.cuda.device(device):
            _lib.check(plan.lib.talfe_detect_padding(audio.data_ptr(), _DTYPES[audio.dtype], audio.shape[0], audio.shape[1],
                                                     audio.stride(0) if audio.shape[0] > 1 else max(audio.stride(0), audio.shape[1]),
                                                     lens.data_ptr(), torch.cuda.current_stream(device).cuda_stream),
                       "talfe_detect_padding")
        return lens

    @torch.jit.ignore
    def features(self, audio: torch.Tensor, audio_lens: Optional[torch.Tensor] = None, norm: str = "batch",
                 layout: str = "tm", out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
                 defer_normalise: bool = False, spec_augment=None, given_stats: Optional[torch.Tensor] = None,
                 lens_are_padding: bool = False) -> torch.Tensor:
        """Extension surface on the same kernels.

        audio_lens  int64 [B] true lengths (what the collaters emit next to the padded batch,
                    tal/asr/data/aligned.py:246-270).  When given, every row is processed as if it were
                    alone: own frame count 1 + len // 160, reflection at its own end, frames beyond
                    written as 0 and excluded from statistics.  Without it rows are taken as padded
                    (the reference's behaviour).
        norm        'batch' (reference) | 'none' | 'row' | 'row_mel' | 'row_mel_var'
        layout      'tm' -> [B, T, M] (reference) | 'mt' -> [B, M, T] (what encode_features wants, models.py:167)
        stats       optional float64 tensor receiving the statistics block(s) (see include/talfe.h)
        spec_augment (freq_bands, time_bands) from ``specaug.sample_masks``: zeroed in the normalisation sweep,
                    replacing ``time_mask(freq_mask(x))`` of tal/asr/models.py:159-161
        lens_are_padding  with ``audio_lens``: keep the REFERENCE semantics (every row padded to L, the padding frames count
                    in the mean) and take ``audio_lens`` as the collater's guarantee that row r is zero from sample
                    ``audio_lens[r]`` on: frames that cannot see a real sample are the constant log(eps) and are filled,
                    not computed — the reference's result at the cost of the real audio (norm 'batch' / 'none')
        given_stats optional float64 statistics block [1, 3 + 2 M] on the device (e.g. ``CorpusStats.block`` after its
                    all-reduce): EVERY row is normalised with it, according to ``norm``, inside the transform kernel —
                    dataset-level CMVN without a sweep over the features; the values are identical to
                    ``features(norm='none')`` followed by ``apply_stats`` with that block
        """
        with torch.no_grad():
            audio = _prepare_audio(audio)
            device = _require_cuda(audio)
            num_frames(audio.shape[1], self.n_fft, self.hop)
            lens = None
            if audio_lens is not None:
                lens = audio_lens.to(device=device, dtype=torch.int64).contiguous()
                if lens.numel() != audio.shape[0]:
                    raise ValueError("audio_lens must have one entry per row")
            bands = None
            if spec_augment is not None:
                if norm == "none" or defer_normalise:
                    raise ValueError("spec_augment rides on the normalisation sweep: needs norm != 'none' and no deferral")
                fb, tb = spec_augment
                fb = fb.to(device=device, dtype=torch.int32).contiguous()
                tb = tb.to(device=device, dtype=torch.int32).contiguous()
                if fb.shape != tb.shape or fb.dim() != 3 or fb.shape[0] != audio.shape[0] or fb.shape[2] != 2 or fb.shape[1] > 16:
                    raise ValueError("spec_augment bands must be two int tensors [B, n_bands <= 16, 2]")
                bands = (fb, tb)
            if lens_are_padding:
                if lens is None or norm not in ("batch", "none") or defer_normalise or spec_augment is not None or given_stats is not None or stats is not None:
                    raise ValueError("lens_are_padding needs audio_lens, norm 'batch' or 'none', and none of stats / defer / spec_augment / given_stats")
            if given_stats is not None:
                if norm == "none" or defer_normalise or spec_augment is not None:
                    raise ValueError("given_stats needs norm != 'none', no deferral and no spec_augment")
                if (given_stats.dtype != torch.float64 or given_stats.device != device or not given_stats.is_contiguous()
                        or given_stats.numel() < _lib.stats_doubles(self.n_mels)):
                    raise ValueError(f"given_stats must be a contiguous float64 block of {_lib.stats_doubles(self.n_mels)} doubles on {device}")
            return _run(self.plan(device), audio, norm=_NORMS[norm], layout=_LAYOUTS[layout], eps=self.eps,
                        lens=lens, out=out, stats=stats, defer=defer_normalise, bands=bands, given_stats=given_stats,
                        padding_hint=lens_are_padding)

    @torch.jit.ignore
    def features_packed(self, audio: torch.Tensor, audio_lens: torch.Tensor, norm: str = "row"):
        """Ragged batch without padding frames: returns (feats [sum T_i, n_mels], frame_offsets
        int64 [B + 1]); row i owns feats[frame_offsets[i]:frame_offsets[i+1]] = the front end run on that row
        alone (T_i = 1 + len_i // 160).  The reference instead pads every row to the longest
        (tal/asr/data/aligned.py:250-257), which makes a 1 s row as expensive as a 10 min one."""
        with torch.no_grad():
            audio = _prepare_audio(audio)
            device = _require_cuda(audio)
            lens_host = audio_lens.detach().to("cpu", torch.int64)
            if lens_host.numel() != audio.shape[0] or int(lens_host.min()) <= self.n_fft // 2 or int(lens_host.max()) > audio.shape[1]:
                raise RuntimeError(f"audio_lens must give every row a length in ({self.n_fft // 2}, L]")
            frames = 1 + (lens_host + 2 * (self.n_fft // 2) - self.n_fft) // self.hop
            offsets_host = torch.zeros(audio.shape[0] + 1, dtype=torch.int64)
            offsets_host[1:] = torch.cumsum(frames, 0)
            offsets = offsets_host.to(device)
            out = _run(self.plan(device), audio, norm=_NORMS[norm], layout=_lib.LAYOUT_TM, eps=self.eps,
                       lens=lens_host.to(device), out_offsets=offsets, packed_frames=int(offsets_host[-1]))
            return out, offsets

    @torch.jit.ignore
    def forward_host(self, audio_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                     device: Optional[torch.device] = None) -> torch.Tensor:
        """The same call for HOST buffers (ideally pinned): H2D of the waveforms, forward on the device,
        D2H of the features into ``out_host``; returns after the copy has completed."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        with torch.no_grad():
            x = audio_host.to(device, non_blocking=True)
            y = self.forward(x)
            if out_host is None:
                out_host = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
            out_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
        return out_host

    def stats_block(self, device, rows: int = 1) -> torch.Tensor:
        return torch.zeros(rows, _lib.stats_doubles(self.n_mels), dtype=torch.float64, device=device)

    def apply_stats(self, feats: torch.Tensor, stats: torch.Tensor, norm: str = "batch", layout: str = "tm",
                    valid_frames: Optional[torch.Tensor] = None) -> torch.Tensor:
        """In-place normalisation from a statistics block (after streaming or a cross-rank all-reduce)."""
        device = _require_cuda(feats)
        if feats.dim() != 3 or feats.dtype != torch.float32 or not feats.is_contiguous():
            raise ValueError("feats must be a contiguous float32 [B, T, M] / [B, M, T] tensor")
        B = feats.shape[0]
        T = feats.shape[1] if layout == "tm" else feats.shape[2]
        plan = self.plan(device)
        vf = valid_frames.to(device=device, dtype=torch.int64).contiguous() if valid_frames is not None else None
        with torch.cuda.device(device):
            _lib.check(plan.lib.talfe_apply_stats(plan.handle, feats.data_ptr(), B, T, 0, _LAYOUTS[layout], _NORMS[norm],
                                                  stats.data_ptr(), vf.data_ptr() if vf is not None else None,
                                                  torch.cuda.current_stream(device).cuda_stream), "talfe_apply_stats")
        return feats
