"""Deterministic synthetic 16 kHz mono audio, integer-only so that the
numpy generator here and the CUDA generator in ``csrc/talfe.cu`` (``talfe_synth_fill``)
produce bit-identical samples for any (seed, episode, sample index) and any sharding.

Signal model (what ``tal/asr/data/util.py:18-53`` hands the front end: int16 PCM / 32768):
broadband noise (sum of four 16-bit uniforms, sigma = 0.1 full scale) shaped by a slow
triangular envelope (0.05 .. 1.0, ~0.3 Hz, per-episode phase) with ~10 % of 0.25 s blocks
forced to exact zero (exercises the ``log(0 + eps)`` floor).
"""
from __future__ import annotations

import numpy as np

ENV_PERIOD = 53333          # samples, ~0.3 Hz at 16 kHz
GAP_BLOCK = 4000            # samples, 0.25 s
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def episode_key(seed: int, episode: int) -> np.uint64:
    with np.errstate(over="ignore"):
        mixed = (np.uint64(seed) ^ (np.uint64(episode) * np.uint64(0xD1B54A32D192ED03))) & _M64
        return _splitmix64(np.array([mixed], dtype=np.uint64))[0]


def pcm16(seed: int, episode: int, start: int, count: int) -> np.ndarray:
    """int16 PCM samples [start, start + count) of one episode."""
    with np.errstate(over="ignore"):
        key = episode_key(seed, episode)
        i = np.arange(start, start + count, dtype=np.uint64)
        h = _splitmix64(key + i)
        s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64) - 131070
        amp = (s * 2838) >> 15                                   # sigma*32767/std(sum of 4 u16)
        eoff = np.uint64(key >> np.uint64(40)) % np.uint64(ENV_PERIOD)
        ph = ((i + eoff) % np.uint64(ENV_PERIOD)).astype(np.int64)
        tri = np.abs(2 * ph - ENV_PERIOD)                        # 0 .. ENV_PERIOD
        env_q16 = 3277 + (62259 * tri) // ENV_PERIOD             # 0.05 .. 1.0 in Q16
        k = (amp * env_q16) >> 16
        block = i // np.uint64(GAP_BLOCK)
        g = _splitmix64(key + np.uint64(0x5851F42D4C957F2D) * (block + np.uint64(1)))
        k = np.where((g >> np.uint64(32)) % np.uint64(10) == 0, 0, k)
        return np.clip(k, -32767, 32767).astype(np.int16)


def waveform(seed: int, episode: int, start: int, count: int) -> np.ndarray:
    """float32 samples in [-1, 1): PCM / 32768 exactly as torchaudio.load normalises."""
    return pcm16(seed, episode, start, count).astype(np.float32) / np.float32(32768.0)


def batch(seed: int, n_rows: int, n_samples: int, first_episode: int = 0) -> np.ndarray:
    """[n_rows, n_samples] float32; row r is episode first_episode + r from sample 0."""
    return np.stack([waveform(seed, first_episode + r, 0, n_samples) for r in range(n_rows)])
