"""WAV slice -> pinned int16 staging: the host-side front of the waveform pipeline.

The reference reads audio with ``torchaudio.load(path, offset, num_frames)`` and gets float32 in [-1, 1]
(``tal/asr/data/util.py:18-53`` ``load_audio_segment``; whole episodes in ``tal/baseline/reconcile.py:76-78``):
the int16 samples on disk (``tal/utils/audio.py:12-13``: 16 kHz, 16-bit, mono) are widened to float32 on the
host, and twice the bytes then cross PCIe.  Here the PCM stays int16 all the way to the kernel, which applies
the same 1/32768 scale inside its window multiply (``TALFE_I16``, include/talfe.h) — bit-for-bit the values
``torchaudio.load`` would have produced, half the host->device traffic, no host-side conversion pass.

    seg = load_audio_segment_pcm16(path, start_s, end_s)           # 1-D int16, pinned; same arguments as the reference
    batch, lens = collate_pcm16([seg0, seg1, ...])                   # zero right-pad + audio_lens (aligned.py:246-270)
    pipe.submit(batch, out_host)                                     # tal_asrd_b200.HostPipeline
    ep = load_episode_pcm16(path)                                    # whole episode for stream_episode (reconcile.py:76)

Only canonical RIFF/WAVE PCM (format 1, 16-bit) and WAVE_FORMAT_EXTENSIBLE with the PCM sub-format are accepted,
which is what the reference's conversion step writes (``tal/utils/audio.py:38``: ffmpeg ``-acodec pcm_s16le -ac 1
-ar 16000``).  Anything else raises: this module has no resampler and no decoder, by design.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

DEFAULT_SR = 16000          # /root/reference/tal/asr/data/util.py:16


@dataclass(frozen=True)
class WavInfo:
    rate: int               # torchaudio.info(...).rate in the reference (util.py:33-34)
    channels: int
    length: int             # sample frames in the data chunk
    data_offset: int        # byte offset of the first sample
    block_align: int


def wav_info(path: str) -> WavInfo:
    """Parses the RIFF header (chunks may come in any order; ``LIST`` etc. are skipped)."""
    with open(path, "rb") as fh:
        head = fh.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            hdr = fh.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                raw = fh.read(size)
                tag, channels, rate, _, block_align, bits = struct.unpack("<HHIIHH", raw[:16])
                if tag == 0xFFFE and size >= 40:                       # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                    tag = struct.unpack("<H", raw[24:26])[0]
                fmt = (tag, channels, rate, block_align, bits)
                if size & 1:
                    fh.seek(1, os.SEEK_CUR)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, channels, rate, block_align, bits = fmt
                if tag != 1 or bits != 16:
                    raise ValueError(f"{path}: only 16-bit PCM is supported (format tag {tag}, {bits} bits); convert as "
                                     f"tal/utils/audio.py does (pcm_s16le)")
                offset = fh.tell()
                avail = os.path.getsize(path) - offset
                if size == 0xFFFFFFFF or size > avail:                 # streamed writers leave the size open
                    size = avail
                return WavInfo(rate, channels, size // block_align, offset, block_align)
            else:
                fh.seek(size + (size & 1), os.SEEK_CUR)


def _pinned_int16(n: int) -> torch.Tensor:
    try:
        return torch.empty(n, dtype=torch.int16, pin_memory=True)
    except RuntimeError:                                                # no CUDA driver (CPU-only test box): pageable
        return torch.empty(n, dtype=torch.int16)


def load_audio_segment_pcm16(audio_path: str, start_s: float = 0.0, end_s: Optional[float] = None,
                             out: Optional[torch.Tensor] = None, sr: int = DEFAULT_SR) -> torch.Tensor:
    """Same slice as the reference's ``load_audio_segment(audio_path, start_s, end_s)`` (util.py:32-41:
    ``offset = int(start_s * rate)``, ``num_frames = int((end_s - start_s) * rate)``, 0 = to the end of the file),
    returned as the file's own int16 samples: 1-D, mono, in pinned host memory (or in ``out[:n]`` when given).
    ``x.float() / 32768`` equals what the reference gets from ``torchaudio.load`` (util.py:43).

    The file must already be at ``sr`` (the reference resamples otherwise, util.py:45-48; this loader raises:
    resampling belongs to the offline conversion step, tal/utils/audio.py)."""
    info = wav_info(audio_path)
    if info.rate != sr:
        raise ValueError(f"{audio_path}: sample rate {info.rate} != {sr}; convert it first (tal/utils/audio.py:38)")
    if info.channels != 1:
        raise ValueError(f"{audio_path}: {info.channels} channels; the corpus is mono (tal/utils/audio.py:13)")
    offset = int(start_s * info.rate)
    duration = end_s - start_s if end_s is not None else 0.0
    frames = int(duration * info.rate)
    if offset < 0 or offset > info.length:
        raise ValueError(f"{audio_path}: offset {offset} outside the file ({info.length} samples)")
    n = info.length - offset if frames <= 0 else min(frames, info.length - offset)
    if out is None:
        out = _pinned_int16(n)
    elif out.dtype != torch.int16 or out.dim() != 1 or out.numel() < n or not out.is_contiguous() or out.is_cuda:
        raise ValueError("out must be a contiguous 1-D int16 host tensor with room for the segment")
    dst = out[:n]
    with open(audio_path, "rb", buffering=0) as fh:
        fh.seek(info.data_offset + 2 * offset)
        got = fh.readinto(memoryview(dst.numpy()).cast("B"))           # straight into the pinned block: no intermediate copy
    if got != 2 * n:
        raise IOError(f"{audio_path}: short read ({got} of {2 * n} bytes)")
    return dst


def load_episode_pcm16(audio_path: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Whole file, like ``torchaudio.load(wav_loc)`` in tal/baseline/reconcile.py:76, as pinned int16."""
    return load_audio_segment_pcm16(audio_path, 0.0, None, out=out)


def collate_pcm16(segments: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The collaters' layout (tal/asr/data/aligned.py:246-270) for int16: zero right-pad to the longest segment,
    ``audio_lens`` = the true lengths.  Returns (pinned int16 [B, Lmax], int64 [B])."""
    lens = torch.tensor([int(s.numel()) for s in segments], dtype=torch.int64)
    B, Lmax = len(segments), int(lens.max())
    if out is None:
        out = _pinned_int16(B * Lmax).view(B, Lmax)
    elif tuple(out.shape) != (B, Lmax) or out.dtype != torch.int16:
        raise ValueError(f"out must be int16 [{B}, {Lmax}]")
    for b, s in enumerate(segments):
        n = int(s.numel())
        out[b, :n] = s
        out[b, n:] = 0
    return out, lens


def write_wav_pcm16(path: str, samples, rate: int = DEFAULT_SR) -> None:
    """Minimal canonical writer (tests and the bench's loader leg): mono, 16-bit PCM."""
    a = np.ascontiguousarray(np.asarray(samples, dtype=np.int16))
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + a.nbytes) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16))
        fh.write(b"data" + struct.pack("<I", a.nbytes))
        fh.write(a.tobytes())


def segments_from_files(items: Sequence[Tuple[str, float, Optional[float]]]) -> List[torch.Tensor]:
    """[(path, start_s, end_s), ...] -> list of pinned int16 segments (one reference ``load_audio_segment`` call each)."""
    return [load_audio_segment_pcm16(p, s, e) for p, s, e in items]
