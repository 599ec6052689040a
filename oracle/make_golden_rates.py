"""TEST INFRASTRUCTURE ONLY — golden vectors for LogMelSpec at sample rates other than 16 kHz.

Runs the UNMODIFIED reference class ``tal.asr.models.LogMelSpec(sr=..., n_mels=...)`` (tal/asr/models.py:15-53; fp32 and
its float64 twin) from /root/reference in the build container and freezes inputs, outputs, buffers and frame counts into
``tests/golden/rates.npz``.  The reference derives n_fft = win = int(0.025 sr) and hop = int(0.010 sr) from ``sr``
(models.py:24-32); only 16 kHz is used by its pipelines, so these cases pin the REST of the constructor's signature.

    python oracle/make_golden_rates.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = [  # (sr, n_mels, samples): even and odd n_fft, hop that does / does not divide the length, fewer mels
    (8000, 80, 6400),
    (8000, 40, 5003),
    (22050, 80, 11025),          # n_fft 551 (odd), hop 220
    (22050, 64, 8800),           # length a multiple of the hop with an odd n_fft: 1 + (L - 1) // hop frames
    (44100, 80, 9000),           # n_fft 1102, hop 441
    (48000, 80, 12000),          # n_fft 1200, hop 480
    (11025, 23, 4000),           # n_fft 275 (odd), hop 110
]


def main():
    if not ref_import.reference_available():
        raise SystemExit("reference tree or torchaudio not available; cannot regenerate golden vectors")
    torch.set_num_threads(1)
    rng = np.random.default_rng(404)
    blob = {"cases": np.array([f"{sr}_{m}_{n}" for sr, m, n in CASES])}
    for sr, n_mels, n in CASES:
        ref32 = ref_import.reference_logmel(double=False, sr=sr, n_mels=n_mels)
        ref64 = ref_import.reference_logmel(double=True, sr=sr, n_mels=n_mels)
        env = 0.05 + 0.95 * np.abs(np.sin(np.arange(n) * (2.0 * np.pi * 1.3 / sr)))
        x = np.round(np.clip(rng.standard_normal((2, n)) * 0.1 * env, -1, 1) * 32767.0) / 32768.0
        x[1, n // 3: n // 3 + n // 10] = 0.0                       # an exact-zero gap (the eps floor)
        x = x.astype(np.float32)
        xt = torch.from_numpy(x)
        y32 = ref32(xt).numpy()
        with torch.no_grad():
            raw64 = torch.log(ref64.mel_transform(xt.double()).permute(0, 2, 1) + ref64.eps).contiguous().numpy()
        y64 = ref64(xt.double()).numpy()
        key = f"{sr}_{n_mels}_{n}"
        blob[key + "_audio"] = x
        blob[key + "_ref_f32"] = y32
        blob[key + "_ref_f64"] = y64
        blob[key + "_ref_f64_unnormalised"] = raw64
        blob[key + "_window"] = ref32.mel_transform.spectrogram.window.numpy()
        blob[key + "_fb"] = ref32.mel_transform.mel_scale.fb.numpy()
        # frame counts around the hop / padding boundaries
        n_fft, hop = int(25 / 1000 * sr), int(10 / 1000 * sr)
        lens = [n_fft // 2 + 1, n_fft, 3 * hop - 1, 3 * hop, 3 * hop + 1, 10 * hop]
        blob[key + "_lens"] = np.array(lens)
        blob[key + "_frames"] = np.array([int(ref32(torch.zeros(1, L)).shape[1]) for L in lens])
        print(f"{key:16s} n_fft {n_fft:5d} hop {hop:4d} out {y32.shape} |f32-f64|max {np.abs(y32 - y64).max():.3e}")
    np.savez_compressed(os.path.join(GOLDEN, "rates.npz"), **blob)


if __name__ == "__main__":
    main()
