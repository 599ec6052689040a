"""TEST INFRASTRUCTURE ONLY — golden vectors for the loader-side resampling: outputs of torchaudio.transforms.Resample
(what tal/asr/data/util.py:44-48 calls) in this container, fp32 and float64, frozen into tests/golden/resample.npz.

    python oracle/make_golden_resample.py
"""
import hashlib
import os

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAIRS = [(44100, 16000, 5003), (48000, 16000, 6000), (8000, 16000, 2500), (22050, 16000, 4410), (11025, 16000, 3001),
         (32000, 16000, 4001), (16000, 16000, 1000)]


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(99)
    blob = {"pairs": np.array(PAIRS)}
    for orig, new, n in PAIRS:
        x = np.round(np.clip(rng.standard_normal((2, n)) * 0.2, -1, 1) * 32767.0) / 32768.0
        x[1, : n // 7] = 0.0
        x = x.astype(np.float32)
        t32 = torchaudio.transforms.Resample(orig_freq=orig, new_freq=new)
        y32 = t32(torch.from_numpy(x)).numpy()
        t64 = torchaudio.transforms.Resample(orig_freq=orig, new_freq=new, dtype=torch.float64)
        y64 = t64(torch.from_numpy(x).double()).numpy()
        key = f"{orig}_{new}"
        blob[key + "_audio"], blob[key + "_ref_f32"], blob[key + "_ref_f64"] = x, y32, y64
        if orig != new:
            k = t32.kernel.numpy()
            blob[key + "_kernel_sha1"] = np.frombuffer(hashlib.sha1(k.tobytes()).digest(), dtype=np.uint8)
            blob[key + "_kernel_shape"] = np.array(k.shape)
        print(f"{key:12s} in {x.shape} out {y32.shape} |f32-f64|max {np.abs(y32 - y64).max():.3e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample.npz"), **blob)


if __name__ == "__main__":
    main()
