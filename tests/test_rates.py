"""LogMelSpec(sr != 16000): the rest of the reference constructor's signature (tal/asr/models.py:22-32 derives
n_fft = win = int(0.025 sr), hop = int(0.010 sr)).  Golden vectors: outputs of the UNMODIFIED reference class at seven
(sr, n_mels) combinations, frozen by oracle/make_golden_rates.py into tests/golden/rates.npz.
CPU part: the oracle's generalised restatement and the host-side tables / frame counts against them.
GPU part (-m gpu): the generic sm_100a kernel (csrc/talfe_generic.cuh) through the C ABI against them."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from oracle import logmel_oracle as O

G = np.load(os.path.join(GOLDEN_DIR, "rates.npz"))
CASES = [str(c) for c in G["cases"]]
TOL = 1e-4


def _case(key):
    sr, n_mels, n = (int(v) for v in key.split("_"))
    return sr, n_mels, {k: G[f"{key}_{k}"] for k in ("audio", "ref_f32", "ref_f64", "ref_f64_unnormalised", "window", "fb", "lens", "frames")}


@pytest.mark.parametrize("key", CASES)
def test_oracle_restatement_matches_the_reference_at_other_rates(key):
    sr, n_mels, c = _case(key)
    window, fb = O.tables_for(sr, n_mels)
    assert np.array_equal(window, c["window"]) and np.array_equal(fb, c["fb"])         # bit-identical buffers
    raw = O.logmel_unnormalised_f64_sr(c["audio"], sr, n_mels)
    assert raw.shape == c["ref_f64_unnormalised"].shape
    assert np.abs(raw - c["ref_f64_unnormalised"]).max() < 1e-9
    assert np.abs(O.logmel_f64_sr(c["audio"], sr, n_mels) - c["ref_f64"]).max() < 1e-9
    n_fft, hop = O.geometry(sr)
    for L, T in zip(c["lens"], c["frames"]):
        assert O.frame_count_general(int(L), n_fft, hop) == int(T)
    with pytest.raises(RuntimeError):
        O.frame_count_general(n_fft // 2, n_fft, hop)


@pytest.mark.parametrize("key", CASES)
def test_host_mirror_tables_and_frame_counts_at_other_rates(key):
    from tal_asrd_b200 import LogMelSpec, num_frames, reference_tables
    sr, n_mels, c = _case(key)
    window, fb = reference_tables(n_mels, sr)
    assert np.array_equal(window.numpy(), c["window"]) and np.array_equal(fb.numpy(), c["fb"])
    m = LogMelSpec(sr=sr, n_mels=n_mels)
    assert (m.n_fft, m.hop) == O.geometry(sr)
    for L, T in zip(c["lens"], c["frames"]):
        assert num_frames(int(L), m.n_fft, m.hop) == int(T)
    with pytest.raises(RuntimeError):
        num_frames(m.n_fft // 2, m.n_fft, m.hop)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("key", CASES)
def test_generic_kernel_matches_the_reference_at_other_rates(dev, key):
    from tal_asrd_b200 import LogMelSpec
    sr, n_mels, c = _case(key)
    mod = LogMelSpec(sr=sr, n_mels=n_mels).to(dev)
    x = torch.from_numpy(c["audio"]).to(dev)
    y = mod(x)
    assert y.dtype == torch.float32 and y.is_contiguous() and tuple(y.shape) == c["ref_f32"].shape     # frame count bit-exact
    got = y.cpu().numpy()
    gap = rel_err(c["ref_f32"], c["ref_f64"])
    assert rel_err(got, c["ref_f64"]) <= max(TOL, 2 * gap), key
    assert rel_err(got, c["ref_f32"]) <= max(TOL, 3 * gap), key
    raw = mod.features(x, norm="none").cpu().numpy()
    assert rel_err(raw, c["ref_f64_unnormalised"]) <= TOL
    # frame counts and the too-short error of the reflect pad
    for L, T in zip(c["lens"], c["frames"]):
        assert mod(torch.zeros(1, int(L), device=dev)).shape == (1, int(T), n_mels)
    with pytest.raises(RuntimeError):
        mod(torch.zeros(1, mod.n_fft // 2, device=dev))


@pytest.mark.gpu
def test_generic_kernel_extension_surface(dev):
    """Per-row lengths, [B, M, T] layout, narrow input types, packed output and the statistics modes run on the same
    later kernels as the 16 kHz path; each is checked against the generalised float64 oracle."""
    from tal_asrd_b200 import LogMelSpec
    sr, n_mels = 22050, 64
    mod = LogMelSpec(sr=sr, n_mels=n_mels).to(dev)
    rng = np.random.default_rng(7)
    L = 3 * 22050 + 17
    x = (np.round(rng.standard_normal((3, L)) * 0.1 * 32767.0) / 32768.0).astype(np.float32)
    lens = np.array([L, 20000, 9001])
    xt = torch.from_numpy(x).to(dev)
    want = [O.logmel_unnormalised_f64_sr(x[i:i + 1, :n], sr, n_mels)[0] for i, n in enumerate(lens)]
    # per-row semantics: own frame count, reflection at the row's own end, zeros beyond, own mean
    y = mod.features(xt, audio_lens=torch.from_numpy(lens), norm="row").cpu().numpy()
    for i, w in enumerate(want):
        assert rel_err(y[i, :w.shape[0]], w - w.mean()) <= TOL
        assert not y[i, w.shape[0]:].any()
    # layout and input types
    full = O.logmel_f64_sr(x, sr, n_mels)
    assert rel_err(mod(xt).cpu().numpy(), full) <= TOL
    assert torch.equal(mod.features(xt, layout="mt"), mod(xt).transpose(1, 2).contiguous())
    pcm = torch.from_numpy(np.round(x * 32768.0).astype(np.int16)).to(dev)
    assert rel_err(mod(pcm).cpu().numpy(), full) <= TOL
    half = xt.half()
    assert rel_err(mod(half).cpu().numpy(), O.logmel_f64_sr(half.float().cpu().numpy(), sr, n_mels)) <= TOL
    # per-mel mean / variance over each row (CMVN) and the packed ragged form
    z = mod.features(xt, audio_lens=torch.from_numpy(lens), norm="row_mel_var").cpu().numpy()
    for i, w in enumerate(want):
        ref = (w - w.mean(0)) / np.sqrt(np.maximum(w.var(0), 1e-10))
        assert np.abs(z[i, :w.shape[0]] - ref).max() <= 2e-3          # (variance of float32 features: looser, as for 16 kHz)
    packed, offsets = mod.features_packed(xt, torch.from_numpy(lens), norm="row")
    off = offsets.cpu().numpy()
    assert off[-1] == sum(w.shape[0] for w in want)
    for i, w in enumerate(want):
        assert rel_err(packed[off[i]:off[i + 1]].cpu().numpy(), w - w.mean()) <= TOL
    # streaming stays a 16 kHz feature
    from tal_asrd_b200.streaming import stream_episode
    with pytest.raises(NotImplementedError):
        stream_episode(mod, xt[0])
