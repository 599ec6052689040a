"""TEST INFRASTRUCTURE ONLY — CPU restatement of the resampling the reference's loader applies to files that are not
16 kHz: ``torchaudio.transforms.Resample(orig_freq=sr, new_freq=16000)`` (tal/asr/data/util.py:44-48), whose algorithm
lives in torchaudio (``functional.py:_get_sinc_resample_kernel`` / ``_apply_sinc_resample_kernel``; torchaudio 0.4.0 is
pinned by the reference's requirements.txt, the container has 2.11: the method — windowed-sinc polyphase FIR, low-pass
width 6 — is the one torchaudio took over from kaldi's LinearResample, which 0.4.0 wraps).
Pinned by tests/golden/resample.npz = outputs of torchaudio itself (oracle/make_golden_resample.py).
Never imported by the product."""
from __future__ import annotations

import math

import numpy as np


def sinc_kernel_f64(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    g = math.gcd(orig_freq, new_freq)
    orig, new = orig_freq // g, new_freq // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return k * window * (base / orig), width, orig, new


def resample_f64(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """[..., L] -> [..., ceil(new L / orig)] float64: zero-pad (width, width + orig), correlate with stride orig."""
    x = np.asarray(x, dtype=np.float64)
    if orig_freq == new_freq:
        return x
    k, width, orig, new = sinc_kernel_f64(orig_freq, new_freq)
    lead = x.shape[:-1]
    rows = x.reshape(-1, x.shape[-1])
    L = rows.shape[1]
    out_len = -(-new * L // orig)
    xp = np.pad(rows, ((0, 0), (width, width + orig)))
    n_blocks = (xp.shape[1] - k.shape[1]) // orig + 1
    frames = np.stack([xp[:, b * orig:b * orig + k.shape[1]] for b in range(n_blocks)], axis=1)     # [rows, blocks, taps]
    y = np.einsum("rbt,jt->rbj", frames, k).reshape(rows.shape[0], -1)[:, :out_len]
    return y.reshape(lead + (out_len,))
