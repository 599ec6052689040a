"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's log-mel path.

Reference being restated (read-only tree, paths relative to /root/reference):

* ``tal/asr/models.py:22-33``  LogMelSpec.__init__  -> MelSpectrogram(sample_rate=sr,
  n_mels, n_fft=int(25/1000*sr), win_length=int(25/1000*sr), hop_length=int(10/1000*sr)),
  eps = 1e-6.
* ``tal/asr/models.py:36-53``  LogMelSpec.forward   -> mel_transform(audio) ->
  permute(0, 2, 1) -> log(mel + eps) -> mel -= mel.mean().

The arithmetic itself lives in a third-party dependency that is NOT vendored in
the reference tree: ``torchaudio==0.4.0`` on ``pytorch 1.4`` (requirements.txt:12,
Dockerfile:1), i.e. ``torchaudio.transforms.MelSpectrogram`` -> ``torch.stft``.
Its published algorithm under the arguments above (all other arguments at their
defaults) is restated here in three independent forms:

1. ``logmel_f64``      numpy float64, explicit reflect framing + ``np.fft.rfft``.
                       The arbiter ("float64 twin") for error budgeting.
2. ``logmel_naive``    pure-Python/numpy direct O(N^2) DFT from the definition,
                       for tiny inputs; shares no FFT code with (1).
3. ``logmel_port_f32`` fp32 torch port of exactly the op sequence the reference
                       executes on CPU (reflect pad, strided frames x Hann, rfft,
                       |.|^2, dense [201x80] mel matmul, log(.+eps), scalar mean).
                       Uses all host threads; this is the timed ``cpu_baseline``.

All three are checked against golden outputs of the real reference class in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import math

import numpy as np

SR = 16000
N_FFT = 400          # int(25/1000 * 16000)   models.py:27
WIN = 400            # int(25/1000 * 16000)   models.py:29
HOP = 160            # int(10/1000 * 16000)   models.py:31
N_FREQ = N_FFT // 2 + 1
N_MELS = 80          # models.py:22
EPS = 1e-6           # models.py:22
LOG_EPS = math.log(EPS)


# --------------------------------------------------------------------------- framing
def frame_count(n_samples: int, hop: int = HOP, n_fft: int = N_FFT) -> int:
    """torch.stft(center=True): T = 1 + L // hop; reflect padding needs L > n_fft // 2.

    models.py:91 comment ("15999 frames => 100 frames") is the reference's only KAT.
    """
    if n_samples <= n_fft // 2:
        raise RuntimeError(
            f"reflect padding of {n_fft // 2} needs more than {n_fft // 2} samples, got {n_samples}")
    return 1 + n_samples // hop


def reflect_index(i, n_samples: int):
    """Index map of F.pad(mode='reflect'): no edge repeat, -k -> k, L-1+k -> L-1-k."""
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    return np.where(i >= n_samples, 2 * (n_samples - 1) - i, i)


def frame_indices(n_samples: int, hop: int = HOP, n_fft: int = N_FFT) -> np.ndarray:
    """[T, n_fft] int64 source-sample index of every element of every frame."""
    t = np.arange(frame_count(n_samples, hop, n_fft), dtype=np.int64)[:, None]
    j = np.arange(n_fft, dtype=np.int64)[None, :]
    return reflect_index(hop * t - n_fft // 2 + j, n_samples)


# --------------------------------------------------------------------------- tables
def hann_periodic(n: int = WIN, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n, periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def mel_filterbank(n_freqs: int = N_FREQ, n_mels: int = N_MELS, sr: int = SR,
                   f_min: float = 0.0, f_max: float | None = None, dtype=np.float64) -> np.ndarray:
    """HTK triangular filters, no area normalisation (torchaudio melscale_fbanks defaults).

    fb[f, m] = max(0, min((f - p_m)/(p_{m+1}-p_m), (p_{m+2} - f)/(p_{m+2}-p_{m+1})))
    with f = linspace(0, sr//2, n_freqs) and p = mel^-1(linspace(mel(f_min), mel(f_max), n_mels+2)).
    """
    if f_max is None:
        f_max = float(sr // 2)
    hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    freqs = np.linspace(0.0, sr // 2, n_freqs, dtype=dtype)
    mel_pts = np.linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2, dtype=dtype)
    hz_pts = (700.0 * (np.power(dtype(10.0), mel_pts / dtype(2595.0)) - 1.0)).astype(dtype)
    width = hz_pts[1:] - hz_pts[:-1]
    delta = hz_pts[None, :] - freqs[:, None]                 # [n_freqs, n_mels + 2]
    rising = -delta[:, :-2] / width[:-1]
    falling = delta[:, 2:] / width[1:]
    return np.maximum(0.0, np.minimum(rising, falling)).astype(dtype)


_TABLES = {}


def reference_tables(n_mels: int = N_MELS):
    """(window[400], fb[201, n_mels]) as float32 numpy arrays, bit-identical to the buffers the
    reference module holds (checked against tests/golden/tables.npz).

    The reference's buffers are computed IN FP32 by torch (``torch.hann_window`` and torchaudio's
    filterbank recipe evaluated with fp32 ``torch.linspace`` / ``pow``), and its ``.double()`` twin
    merely widens those fp32 values.  The analytic float64 formulas above differ from them by up to
    2.4e-7 (window) and 7e-6 (filterbank), so every oracle path uses these tables; torch's fp32
    elementwise kernels are the only way to get the same roundings.
    """
    if n_mels not in _TABLES:
        import torch
        window = torch.hann_window(WIN, periodic=True, dtype=torch.float32)
        freqs = torch.linspace(0, SR // 2, N_FREQ)
        top = 2595.0 * math.log10(1.0 + (SR // 2) / 700.0)
        mel_pts = torch.linspace(0.0, top, n_mels + 2)
        hz_pts = 700.0 * (10.0 ** (mel_pts / 2595.0) - 1.0)
        width = hz_pts[1:] - hz_pts[:-1]
        delta = hz_pts.unsqueeze(0) - freqs.unsqueeze(1)
        rising = (-1.0 * delta[:, :-2]) / width[:-1]
        falling = delta[:, 2:] / width[1:]
        fb = torch.clamp(torch.minimum(rising, falling), min=0.0)
        _TABLES[n_mels] = (window.numpy().copy(), fb.numpy().copy())
    return _TABLES[n_mels]


# --------------------------------------------------------------------------- other sample rates (models.py:24-32)
def geometry(sr: int):
    """(n_fft, hop) = (int(0.025 sr), int(0.010 sr)): what LogMelSpec.__init__ hands to MelSpectrogram."""
    return int(25 / 1000 * sr), int(10 / 1000 * sr)


def frame_count_general(n_samples: int, n_fft: int, hop: int) -> int:
    """torch.stft(center=True): pad n_fft // 2 each side, frames of n_fft every hop -> 1 + (L + 2 (n_fft // 2) - n_fft) // hop
    (= 1 + L // hop for even n_fft, 1 + (L - 1) // hop for odd)."""
    if n_samples <= n_fft // 2:
        raise RuntimeError(f"reflect padding of {n_fft // 2} needs more than {n_fft // 2} samples, got {n_samples}")
    return 1 + (n_samples + 2 * (n_fft // 2) - n_fft) // hop


_TABLES_SR = {}


def tables_for(sr: int, n_mels: int = N_MELS):
    """reference_tables for any sample rate: the module's fp32 buffers (window[n_fft], fb[n_fft // 2 + 1, n_mels])."""
    key = (sr, n_mels)
    if key not in _TABLES_SR:
        import torch
        n_fft, _ = geometry(sr)
        window = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
        freqs = torch.linspace(0, sr // 2, n_fft // 2 + 1)
        top = 2595.0 * math.log10(1.0 + (sr // 2) / 700.0)
        mel_pts = torch.linspace(0.0, top, n_mels + 2)
        hz_pts = 700.0 * (10.0 ** (mel_pts / 2595.0) - 1.0)
        width = hz_pts[1:] - hz_pts[:-1]
        delta = hz_pts.unsqueeze(0) - freqs.unsqueeze(1)
        fb = torch.clamp(torch.minimum((-1.0 * delta[:, :-2]) / width[:-1], delta[:, 2:] / width[1:]), min=0.0)
        _TABLES_SR[key] = (window.numpy().copy(), fb.numpy().copy())
    return _TABLES_SR[key]


def logmel_unnormalised_f64_sr(audio: np.ndarray, sr: int, n_mels: int = N_MELS, eps: float = EPS) -> np.ndarray:
    """[B, L] -> [B, T, n_mels] float64 for LogMelSpec(sr, n_mels) before the mean subtraction: explicit framing
    (reflect, frame t = padded[hop t, hop t + n_fft)), periodic Hann, rfft, |.|^2, filterbank, log(. + eps)."""
    audio = np.asarray(audio, dtype=np.float64)
    n_fft, hop = geometry(sr)
    window, fb = tables_for(sr, n_mels)
    out = []
    for row in audio:
        T = frame_count_general(row.shape[-1], n_fft, hop)
        t = np.arange(T, dtype=np.int64)[:, None]
        j = np.arange(n_fft, dtype=np.int64)[None, :]
        frames = row[reflect_index(hop * t - n_fft // 2 + j, row.shape[-1])] * window.astype(np.float64)[None, :]
        spec = np.fft.rfft(frames, n=n_fft, axis=-1)
        out.append(np.log((spec.real ** 2 + spec.imag ** 2) @ fb.astype(np.float64) + eps))
    return np.stack(out)


def logmel_f64_sr(audio: np.ndarray, sr: int, n_mels: int = N_MELS, eps: float = EPS) -> np.ndarray:
    y = logmel_unnormalised_f64_sr(audio, sr, n_mels, eps)
    return y - y.mean()


# --------------------------------------------------------------------------- float64 twin
def power_frames_f64(x: np.ndarray) -> np.ndarray:
    """[L] -> [T, 201] float64 power spectrum of Hann-windowed, centre/reflect frames."""
    x = np.asarray(x, dtype=np.float64)
    frames = x[frame_indices(x.shape[-1])] * reference_tables()[0].astype(np.float64)[None, :]
    spec = np.fft.rfft(frames, n=N_FFT, axis=-1)
    return spec.real ** 2 + spec.imag ** 2


def logmel_unnormalised_f64(audio: np.ndarray, n_mels: int = N_MELS, eps: float = EPS) -> np.ndarray:
    """[B, L] -> [B, T, n_mels] float64, log(mel + eps) before the mean subtraction."""
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim != 2:
        raise ValueError("audio must be [batch, samples]")
    fb = reference_tables(n_mels)[1].astype(np.float64)
    return np.stack([np.log(power_frames_f64(row) @ fb + eps) for row in audio])


def logmel_f64(audio: np.ndarray, n_mels: int = N_MELS, eps: float = EPS) -> np.ndarray:
    """Reference semantics (models.py:50-52): subtract ONE scalar mean over the whole batch."""
    y = logmel_unnormalised_f64(audio, n_mels, eps)
    return y - y.mean()


# --------------------------------------------------------------------------- definition-level check
def logmel_naive(audio: np.ndarray, n_mels: int = N_MELS, eps: float = EPS) -> np.ndarray:
    """Direct O(N^2) DFT from the definition (tiny inputs only). Shares no FFT with the above."""
    audio = np.asarray(audio, dtype=np.float64)
    n = np.arange(N_FFT)
    k = np.arange(N_FREQ)
    ang = -2.0 * np.pi * np.outer(n, k) / N_FFT
    cos_t, sin_t = np.cos(ang), np.sin(ang)
    win = reference_tables(n_mels)[0].astype(np.float64)
    fb = reference_tables(n_mels)[1].astype(np.float64)
    out = []
    for row in audio:
        L = row.shape[0]
        T = frame_count(L)
        feats = np.empty((T, n_mels))
        for t in range(T):
            seg = np.empty(N_FFT)
            for j in range(N_FFT):
                i = HOP * t - N_FFT // 2 + j
                if i < 0:
                    i = -i
                if i >= L:
                    i = 2 * (L - 1) - i
                seg[j] = row[i] * win[j]
            re, im = seg @ cos_t, seg @ sin_t
            feats[t] = np.log((re * re + im * im) @ fb + eps)
        out.append(feats)
    y = np.stack(out)
    return y - y.mean()


# --------------------------------------------------------------------------- extension semantics (float64)
def normalise_f64(y: np.ndarray, mode: str, lens_frames=None) -> np.ndarray:
    """Normalisation variants over un-normalised log-mel y[B, T, M] (float64).

    'none'            identity
    'batch'           reference: one scalar mean over everything (models.py:52)
    'row'             per-row scalar mean over that row's valid frames
    'row_mel'         per-row, per-mel mean (CMN)
    'row_mel_var'     per-row, per-mel mean and population std (CMVN), std floor 1e-5 on variance
    Frames t >= lens_frames[b] are excluded from the statistics and set to 0.
    """
    y = np.array(y, dtype=np.float64, copy=True)
    B, T, M = y.shape
    if lens_frames is None:
        lens_frames = [T] * B
    if mode == "none":
        for b in range(B):
            y[b, lens_frames[b]:] = 0.0
        return y
    if mode == "batch":
        total = sum(y[b, :lens_frames[b]].sum() for b in range(B))
        count = sum(lens_frames[b] * M for b in range(B))
        mu = total / count
        for b in range(B):
            y[b, :lens_frames[b]] -= mu
            y[b, lens_frames[b]:] = 0.0
        return y
    for b in range(B):
        v = y[b, :lens_frames[b]]
        if mode == "row":
            v -= v.mean()
        elif mode == "row_mel":
            v -= v.mean(axis=0, keepdims=True)
        elif mode == "row_mel_var":
            mu = v.mean(axis=0, keepdims=True)
            var = np.maximum(((v - mu) ** 2).mean(axis=0, keepdims=True), 1e-10)
            v[:] = (v - mu) / np.sqrt(var)
        else:
            raise ValueError(mode)
        y[b, lens_frames[b]:] = 0.0
    return y


def logmel_rows_f64(rows, n_mels: int = N_MELS, eps: float = EPS, mode: str = "row"):
    """Each row processed as if alone (the reference run with B=1 per row), own frame count,
    reflect at the row's own end; returns ([B, Tmax, M] zero-padded, frames per row)."""
    feats = [logmel_unnormalised_f64(np.asarray(r)[None, :], n_mels, eps)[0] for r in rows]
    Tmax = max(f.shape[0] for f in feats)
    y = np.zeros((len(feats), Tmax, n_mels))
    for b, f in enumerate(feats):
        y[b, :f.shape[0]] = f
    lens = [f.shape[0] for f in feats]
    return normalise_f64(y, mode, lens), lens


# --------------------------------------------------------------------------- fp32 port (timed CPU baseline)
_PORT_CACHE = {}


def _port_tables(n_mels: int):
    import torch
    if n_mels not in _PORT_CACHE:
        window, fb = reference_tables(n_mels)
        _PORT_CACHE[n_mels] = (torch.from_numpy(window), torch.from_numpy(fb))
    return _PORT_CACHE[n_mels]


def logmel_port_f32(audio, n_mels: int = N_MELS, eps: float = EPS, normalise: bool = True):
    """fp32 torch port of the op sequence the reference runs on CPU.

    models.py:45  mel_transform(audio): torchaudio Spectrogram (torch.stft, center, reflect,
                  periodic Hann, onesided, power=2) then MelScale (matmul with fb[201, n_mels]);
    models.py:48  permute(0, 2, 1);  models.py:50  log(mel + eps);  models.py:52  mel -= mel.mean().
    Runs on however many threads torch is configured with (all host cores by default).
    """
    import torch
    audio = torch.as_tensor(audio)
    if audio.dim() != 2:
        raise ValueError("audio must be [batch, samples]")
    frame_count(audio.shape[-1])                      # same error as the reference for short input
    window, fb = _port_tables(n_mels)
    with torch.no_grad():
        spec = torch.stft(audio.float(), n_fft=N_FFT, hop_length=HOP, win_length=WIN, window=window,
                          center=True, pad_mode="reflect", normalized=False, onesided=True,
                          return_complex=True)                      # [B, 201, T]
        power = spec.abs().pow(2.0)
        mel = torch.matmul(power.transpose(-1, -2), fb)           # [B, T, n_mels]
        mel = torch.log(mel + eps)
        if normalise:
            mel -= mel.mean()
    return mel.contiguous()
