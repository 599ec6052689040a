"""Loader-side resampling (tal/asr/data/util.py:44-48: torchaudio.transforms.Resample(sr, 16000) for files that are not
16 kHz).  Golden vectors: outputs of torchaudio itself, fp32 and float64 (oracle/make_golden_resample.py).
CPU: the oracle's float64 restatement and the binding's filter table.  GPU (-m gpu): the CUDA kernel through the C ABI."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import resample_oracle as R

G = np.load(os.path.join(GOLDEN_DIR, "resample.npz"))
PAIRS = [tuple(int(v) for v in p) for p in G["pairs"]]


@pytest.mark.parametrize("orig,new,n", PAIRS)
def test_oracle_restatement_matches_torchaudio(orig, new, n):
    key = f"{orig}_{new}"
    y = R.resample_f64(G[key + "_audio"], orig, new)
    assert y.shape == G[key + "_ref_f64"].shape == (2, -(-new * n // orig))          # ceil(new L / orig)
    assert np.abs(y - G[key + "_ref_f64"]).max() < 1e-12


@pytest.mark.parametrize("orig,new,n", [p for p in PAIRS if p[0] != p[1]])
def test_filter_table_is_torchaudios(orig, new, n):
    from tal_asrd_b200.resample import sinc_resample_kernel
    key = f"{orig}_{new}"
    k, width, o, m = sinc_resample_kernel(orig, new)
    assert tuple(G[key + "_kernel_shape"]) == (m, 1, 2 * width + o)
    assert hashlib.sha1(k.numpy().tobytes()).digest() == G[key + "_kernel_sha1"].tobytes()      # bit-identical
    # (torchaudio evaluates j / new in float32 on this path: its table is up to ~1e-5 from the float64 formula)
    assert np.abs(k.numpy() - R.sinc_kernel_f64(orig, new)[0]).max() < 2e-5


def test_module_contract_without_gpu():
    from tal_asrd_b200.resample import Resample
    x = torch.zeros(2, 100)
    assert Resample(16000, 16000)(x) is x                                              # equal rates: unchanged, like torchaudio
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        Resample(44100, 16000)(x)
    assert not list(Resample(44100, 16000).state_dict())                               # the table is not checkpoint state


@pytest.mark.gpu
@pytest.mark.parametrize("orig,new,n", PAIRS)
def test_gpu_resampler_matches_torchaudio(orig, new, n):
    from tal_asrd_b200.resample import Resample
    dev = torch.device("cuda:0")
    key = f"{orig}_{new}"
    x = torch.from_numpy(G[key + "_audio"]).to(dev)
    y = Resample(orig, new).to(dev)(x)
    torch.cuda.synchronize()
    ref32, ref64 = G[key + "_ref_f32"], G[key + "_ref_f64"]
    assert y.dtype == torch.float32 and tuple(y.shape) == ref32.shape
    got = y.cpu().numpy()
    gap = np.abs(ref32 - ref64).max()
    assert np.abs(got - ref32).max() <= 2e-6                                           # same table, other summation order
    assert np.abs(got - ref64).max() <= max(2e-6, 2 * gap)
    if orig != new:
        # int16 PCM in (scaled like torchaudio.load), leading dimensions kept, and straight into the front end
        pcm = torch.from_numpy(np.round(G[key + "_audio"] * 32768.0).astype(np.int16)).to(dev)
        assert np.abs(Resample(orig, new).to(dev)(pcm).cpu().numpy() - ref32).max() <= 2e-6
        assert tuple(Resample(orig, new).to(dev)(x[None]).shape) == (1,) + ref32.shape
        from tal_asrd_b200 import LogMelSpec
        from oracle import logmel_oracle as O
        feats = LogMelSpec().to(dev)(y)
        want = O.logmel_f64(ref64)
        assert feats.shape == want.shape
        if orig > new:          # (after UPsampling the band above the old Nyquist holds ~1e-10 of power: its log is ill-conditioned)
            assert np.abs(feats.cpu().numpy() - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
