"""Resampling to the model's rate on the GPU: the loader-side neighbour of the front end.

The reference's ``load_audio_segment`` (``tal/asr/data/util.py:44-48``) resamples a file whose rate is not 16 kHz with
``torchaudio.transforms.Resample(orig_freq=sr, new_freq=16000)`` on the host, one file at a time.  ``Resample`` here has
that constructor and call signature; the waveform stays on the device (``wavio`` stages raw PCM, the copy engine moves it,
this kernel resamples it, ``LogMelSpec`` consumes it), and int16 PCM is accepted directly (scaled by 1/32768 like
``torchaudio.load``).  The filter is torchaudio's windowed-sinc polyphase FIR (``sinc_interp_hann``, low-pass width 6,
roll-off 0.99; ``torchaudio/functional/functional.py:_get_sinc_resample_kernel``), restated below with the same torch
operations in the same order so that the table is the one torchaudio would build.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.int16: _lib.I16}


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """(kernel [new, 2 width + orig] float32, width, orig, new) with orig / new reduced by their gcd."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
        raise ValueError("frequencies must be integers")
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive.")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=None)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(dtype=torch.float32).reshape(new, 2 * width + orig).contiguous(), width, orig, new


class Resample(nn.Module):
    """Drop-in for ``torchaudio.transforms.Resample(orig_freq, new_freq)`` as the reference uses it
    (tal/asr/data/util.py:45-47) for waveforms that live on a CUDA device.

    forward(waveform[..., L]) -> float32 [..., ceil(new_freq * L / orig_freq)]; float32 / float16 / int16 input.
    Equal rates return the waveform unchanged, like torchaudio."""

    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000, lowpass_filter_width: int = 6, rolloff: float = 0.99):
        super().__init__()
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.lowpass_filter_width, self.rolloff = lowpass_filter_width, rolloff
        if self.orig_freq != self.new_freq:
            k, self.width, self._orig, self._new = sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width, rolloff)
            self.register_buffer("kernel", k, persistent=False)

    @torch.jit.ignore
    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.orig_freq == self.new_freq:
            return waveform
        if not waveform.is_cuda:
            raise RuntimeError("tal_asrd_b200.Resample runs only on a CUDA device (sm_100a); there is no CPU implementation")
        if waveform.dtype not in _DTYPES:
            raise TypeError(f"Expected float32 / float16 / int16 waveform, but received {waveform.dtype}.")
        with torch.no_grad():
            shape = waveform.shape
            x = waveform.reshape(-1, shape[-1])
            if x.stride(-1) != 1:
                x = x.contiguous()
            B, L = x.shape
            out_len = (self._new * L + self._orig - 1) // self._orig
            kernel = self.kernel if self.kernel.device == x.device else self.kernel.to(x.device)
            out = torch.empty(B, out_len, dtype=torch.float32, device=x.device)
            lib = _lib.load()
            with torch.cuda.device(x.device):
                for b0 in range(0, B, 65535):
                    xb, ob = x[b0:b0 + 65535], out[b0:b0 + 65535]
                    _lib.check(lib.talfe_resample(xb.data_ptr(), _DTYPES[x.dtype], xb.shape[0], L, xb.stride(0) if xb.shape[0] > 1 else max(xb.stride(0), L),
                                                  self._orig, self._new, self.width, kernel.data_ptr(), ob.data_ptr(), out_len, out_len,
                                                  torch.cuda.current_stream(x.device).cuda_stream), "talfe_resample")
            return out.reshape(shape[:-1] + (out_len,))
