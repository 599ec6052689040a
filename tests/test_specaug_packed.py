"""SpecAugment masks fused into the normalisation sweep, packed ragged output."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err


def _golden_masks():
    z = np.load(os.path.join(GOLDEN_DIR, "specaug_masks.npz"))
    shape = tuple(int(v) for v in z["shape"])
    n = int(np.prod(shape))
    return shape, {int(k.split("_")[1]): np.unpackbits(z[k])[:n].reshape(shape).astype(bool) for k in z.files if k.startswith("seed_")}


def test_sample_masks_reproduces_reference_random_sequence():
    """Same `random` seed -> the bands the reference's time_mask(freq_mask(x)) zeroes (frozen from the real functions)."""
    from tal_asrd_b200 import specaug
    shape, golden = _golden_masks()
    for seed, want in golden.items():
        random.seed(seed)
        fb, tb = specaug.sample_masks(shape[0], shape[1], shape[2])
        got = specaug.apply_masks_reference(torch.ones(shape), fb, tb) == 0
        assert np.array_equal(got.numpy(), want), seed
        assert fb.dtype == torch.int32 and fb.shape == (shape[0], 2, 2) and tb.shape == (shape[0], 2, 2)


def test_sample_masks_against_live_reference_functions():
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    from tal_asrd_b200 import specaug
    ref = ref_import.load_reference_models()
    for seed in range(40):
        random.seed(seed)
        want = ref.time_mask(ref.freq_mask(torch.ones(3, 250, 80)))
        random.seed(seed)
        fb, tb = specaug.sample_masks(3, 250, 80)
        assert torch.equal(specaug.apply_masks_reference(torch.ones(3, 250, 80), fb, tb), want), seed


def test_encoder_padding_mask_helper():
    """models.py:178-187 builds the mask with a host loop over rows; the helper is the vectorised equivalent."""
    from tal_asrd_b200.frontend import encoder_padding_mask
    lens = torch.tensor([480000, 16000, 250000, 479999])
    enc_T = 358
    scaled = lens // (lens.max() // enc_T)
    want = torch.zeros(4, enc_T, dtype=torch.bool)
    for i, l in enumerate(scaled.tolist()):
        want[i, l:] = 1
    assert torch.equal(encoder_padding_mask(lens, enc_T), want)


@pytest.mark.gpu
def test_specaug_fused_equals_reference_application():
    from tal_asrd_b200 import LogMelSpec, specaug, synth
    dev = torch.device("cuda:0")
    mod = LogMelSpec().to(dev)
    x = torch.from_numpy(synth.batch(3, 4, 48000)).to(dev)
    plain = mod(x)
    for seed in (0, 1, 2, 7):
        random.seed(seed)
        fb, tb = specaug.sample_masks(4, plain.shape[1], 80)
        fused = mod.features(x, spec_augment=(fb, tb))
        want = specaug.apply_masks_reference(plain.cpu(), fb, tb)
        assert torch.equal(fused.cpu() == 0, want == 0)                 # exactly the same cells are zeroed
        assert float((fused.cpu() - want).abs().max()) < 1e-6
    fused_mt = mod.features(x, spec_augment=(fb, tb), layout="mt")
    assert torch.equal(fused_mt.transpose(1, 2).contiguous(), fused)
    with pytest.raises(ValueError):
        mod.features(x, norm="none", spec_augment=(fb, tb))


@pytest.mark.gpu
def test_packed_ragged_output():
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import LogMelSpec, synth
    dev = torch.device("cuda:0")
    mod = LogMelSpec().to(dev)
    lens = [16000, 7777, 48000, 201, 30000]
    Lmax = max(lens)
    x = np.zeros((len(lens), Lmax), np.float32)
    rows = [synth.waveform(5, i, 0, n) for i, n in enumerate(lens)]
    for i, r in enumerate(rows):
        x[i, :len(r)] = r
    xt = torch.from_numpy(x).to(dev)
    for mode in ("row", "none", "batch", "row_mel"):
        packed, off = mod.features_packed(xt, torch.tensor(lens), norm=mode)
        padded = mod.features(xt, audio_lens=torch.tensor(lens), norm=mode)
        torch.cuda.synchronize()
        off = off.cpu().tolist()
        assert off[0] == 0 and off[-1] == packed.shape[0] == sum(1 + n // 160 for n in lens)
        for i, n in enumerate(lens):
            T = 1 + n // 160
            assert torch.equal(packed[off[i]:off[i + 1]], padded[i, :T]), (mode, i)   # same kernel, same values
    ref, frames = O.logmel_rows_f64(rows, mode="row")
    packed, off = mod.features_packed(xt, torch.tensor(lens), norm="row")
    off = off.cpu().tolist()
    for i in range(len(lens)):
        assert rel_err(packed[off[i]:off[i + 1]].cpu().numpy(), ref[i, :frames[i]]) < 1e-4
    with pytest.raises(RuntimeError):
        mod.features_packed(xt, torch.tensor([16000, 100, 48000, 201, 30000]))
