"""SpecAugment masks for the fused normalisation sweep.

The reference applies ``time_mask(freq_mask(x))`` to the features during training
(``tal/asr/models.py:159-161``; ``freq_mask`` :531-548, ``time_mask`` :550-566): per batch row and per
mask it draws a band with Python's global ``random`` and sets it to 0 on a clone, in two more passes over
the ``[B, T, 80]`` tensor plus per-row Python loops.  Here the bands are drawn on the host with exactly
the reference's sequence of ``random`` calls (same seed -> same bands, including its early-return quirk
when a zero width is drawn) and handed to the kernel that normalises the features, which zeroes them in
the same sweep: no extra pass over the features.

    freq, time = sample_masks(B, T, 80)                      # replaces the two reference functions
    x = logmelspec.features(audio, spec_augment=(freq, time))
"""
from __future__ import annotations

import random as _random
from typing import Optional, Tuple

import torch


def _bands(batch: int, extent: int, width: int, num_masks: int, rng) -> list:
    """One axis of the reference's masking loop.  Returns per row a list of (start, end) with
    start == end for "no mask"; mirrors the control flow of freq_mask / time_mask line by line:
        w = randrange(0, width); z = randrange(0, extent - w)
        if z == z + w: return          # zero width drawn: the WHOLE function returns, later rows stay unmasked
        end = randrange(z, z + w); x[b, z:end] = 0
    """
    out = [[(0, 0)] * num_masks for _ in range(batch)]
    for b in range(batch):
        for k in range(num_masks):
            w = rng.randrange(0, width)
            z = rng.randrange(0, extent - w)
            if z == z + w:
                return out
            end = rng.randrange(z, z + w)
            out[b][k] = (z, end)
    return out


def sample_masks(batch: int, n_frames: int, n_mels: int = 80, F: int = 27, T: int = 100, num_masks: int = 2,
                 rng=None, device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(freq_bands, time_bands): int32 tensors [batch, num_masks, 2] of (first, end) pairs, drawn with the same
    calls, in the same order, as ``time_mask(freq_mask(x))`` makes on ``random`` (pass ``rng=random`` or a
    ``random.Random`` instance; default: the global module, like the reference)."""
    rng = rng or _random
    freq = _bands(batch, n_mels, F, num_masks, rng)          # freq_mask runs first (it is the inner call)
    time = _bands(batch, n_frames, T, num_masks, rng)
    f = torch.tensor(freq, dtype=torch.int32).reshape(batch, num_masks, 2)
    t = torch.tensor(time, dtype=torch.int32).reshape(batch, num_masks, 2)
    if device is not None:
        f, t = f.to(device), t.to(device)
    return f, t


def apply_masks_reference(x: torch.Tensor, freq_bands: torch.Tensor, time_bands: torch.Tensor) -> torch.Tensor:
    """Plain-torch application of the bands (host-side helper for tests and for CPU tensors)."""
    y = x.clone()
    for b in range(y.shape[0]):
        for lo, hi in freq_bands[b].tolist():
            y[b, :, lo:hi] = 0
        for lo, hi in time_bands[b].tolist():
            y[b, lo:hi, :] = 0
    return y
