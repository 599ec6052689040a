"""TEST INFRASTRUCTURE ONLY — freeze outputs of the UNMODIFIED reference into tests/golden/.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

For every case it stores the input waveform, the fp32 output of the reference's own
``tal.asr.models.LogMelSpec`` (tal/asr/models.py:15-53) and the output of its float64 twin
(the same module after ``.double()``), plus the module's two buffers (Hann window, mel
filterbank).  The fixtures are small (< 2 MB total) and committed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from tal_asrd_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def lcg_noise(n: int, seed: int = 12345) -> np.ndarray:
    """SURVEY.md §8c vector (4): s <- (1103515245 s + 12345) mod 2^31, x = 0.25 (2 s / 2^31 - 1)."""
    out = np.empty(n, dtype=np.float64)
    s = seed
    for i in range(n):
        s = (1103515245 * s + 12345) % (1 << 31)
        out[i] = 0.25 * (2.0 * s / float(1 << 31) - 1.0)
    return out.astype(np.float32)


def cases():
    n = np.arange(16000)
    rng = np.random.default_rng(2020)
    yield "zeros", np.zeros((1, 3200), np.float32)
    yield "dc_half", np.full((1, 3200), 0.5, np.float32)
    yield "tone_1k", (0.5 * np.cos(2 * np.pi * 1000.0 * n / 16000.0)).astype(np.float32)[None, :]
    yield "lcg_noise", lcg_noise(16000)[None, :]
    yield "synth_batch3", synth.batch(2020, 3, 24000)                 # envelope + exact-zero gaps
    yield "gauss_batch2_ragged_pad", np.concatenate(                  # collater-style zero right-pad
        [rng.standard_normal((1, 12000)).astype(np.float32) * 0.1,
         np.concatenate([rng.standard_normal((1, 7777)).astype(np.float32) * 0.1,
                         np.zeros((1, 12000 - 7777), np.float32)], axis=1)], axis=0)
    for L in (201, 400, 15999, 16000, 16001):                          # edge lengths (models.py:91 KAT)
        yield f"len_{L}", synth.waveform(7, L, 0, L)[None, :]
    yield "loud_fullscale", (rng.uniform(-1, 1, (1, 4000)).astype(np.float32))
    yield "tiny_amplitude", (rng.standard_normal((1, 4000)).astype(np.float32) * 1e-4)


def main():
    if not ref_import.reference_available():
        raise SystemExit("reference tree or torchaudio not available; cannot regenerate golden vectors")
    os.makedirs(GOLDEN, exist_ok=True)
    ref32 = ref_import.reference_logmel(double=False)
    ref64 = ref_import.reference_logmel(double=True)
    torch.set_num_threads(1)            # fixed summation order inside torch for reproducible fixtures

    np.savez_compressed(
        os.path.join(GOLDEN, "tables.npz"),
        window=ref32.mel_transform.spectrogram.window.numpy(),
        fb=ref32.mel_transform.mel_scale.fb.numpy(),
        window64=ref64.mel_transform.spectrogram.window.numpy(),
        fb64=ref64.mel_transform.mel_scale.fb.numpy(),
    )
    names = []
    for name, x in cases():
        xt = torch.from_numpy(x)
        y32 = ref32(xt).numpy()
        x64 = xt.double()
        # float64 twin, before and after the scalar-mean subtraction (forward() subtracts in place)
        with torch.no_grad():
            mel64 = ref64.mel_transform(x64).permute(0, 2, 1)
            raw64 = torch.log(mel64 + ref64.eps).contiguous().numpy()
        y64 = ref64(x64).numpy()
        assert y32.shape == (x.shape[0], 1 + x.shape[1] // 160, 80), y32.shape
        np.savez_compressed(os.path.join(GOLDEN, f"case_{name}.npz"),
                            audio=x, ref_f32=y32, ref_f64=y64, ref_f64_unnormalised=raw64)
        names.append(name)
        print(f"{name:28s} in {x.shape} out {y32.shape} |f32-f64|max {np.abs(y32 - y64).max():.3e}")

    # frame-count table (reference returns these shapes; short inputs raise inside F.pad)
    counts = {}
    for L in (201, 320, 400, 15999, 16000, 16001, 480000, 960000):
        counts[L] = int(ref32(torch.zeros(1, L)).shape[1])
    errors = []
    for L in (1, 160, 200):
        try:
            ref32(torch.zeros(1, L))
        except RuntimeError:
            errors.append(L)
    np.savez_compressed(os.path.join(GOLDEN, "frame_counts.npz"),
                        lengths=np.array(list(counts.keys())), frames=np.array(list(counts.values())),
                        raises_runtime_error=np.array(errors))
    # SpecAugment (tal/asr/models.py:531-566): what time_mask(freq_mask(x)) zeroes for a given `random` seed
    import random
    ref_models = ref_import.load_reference_models()
    sa = {}
    for seed in (0, 1, 2, 3, 7, 11, 2020):
        random.seed(seed)
        ones = torch.ones(4, 300, 80)
        masked = ref_models.time_mask(ref_models.freq_mask(ones))
        sa[f"seed_{seed}"] = np.packbits((masked == 0).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "specaug_masks.npz"), shape=np.array([4, 300, 80]), **sa)
    with open(os.path.join(GOLDEN, "CASES.txt"), "w") as fh:
        fh.write("\n".join(names) + "\n")
    print("frame counts", counts, "errors", errors)


if __name__ == "__main__":
    main()
