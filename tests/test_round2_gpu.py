"""Round-2 GPU tests (through the C ABI like the rest of the GPU suite): the fused in-kernel normalisation, the
drop-in contract under its real caller (extract_features body, half-precision module, deepcopy / pickle), and
full-size parity against the float64 oracle for BASELINE configs 2, 3 and 4 (every row / random windows along
the hour instead of a few samples)."""
import copy
import io
import os
import pickle
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _module(dev, **env):
    """A LogMelSpec whose plan is created under the given development switches (read at plan creation)."""
    from tal_asrd_b200 import LogMelSpec
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        m = LogMelSpec().to(dev)
        m.plan(dev)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return m


def _fill(dev, rows, n, episode=0, dtype=torch.float32):
    from tal_asrd_b200 import _lib
    lib = _lib.load()
    code = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.int16: _lib.I16}[dtype]
    x = torch.empty(rows, n, dtype=dtype, device=dev)
    _lib.check(lib.talfe_synth_fill(x.data_ptr(), code, rows, n, n, 2020, episode, 0, None))
    return x


# ------------------------------------------------------------------------------------------ tensor-copy waveform tiles
def test_tensor_copy_tiles_are_bitwise_the_bulk_copy_path(dev):
    """fp32 interior tiles arrive by ONE cp.async.bulk.tensor (overlapping-row tensor map) instead of 17 bulk pieces:
    same bytes in shared memory, hence identical features — batch, one long row (tiles cross multiples of 256 tensor
    rows), a chunk in the middle of an episode (frame0 / origin) and a strided batch."""
    tma, bulk = _module(dev, TALFE_TMA=1), _module(dev, TALFE_TMA=0)
    x = _fill(dev, 5, 16000 * 21 + 77, episode=3)
    assert torch.equal(tma.features(x, norm="none"), bulk.features(x, norm="none"))
    assert torch.equal(tma(x), bulk(x))
    long_row = _fill(dev, 1, 16000 * 400, episode=4)
    assert torch.equal(tma.features(long_row, norm="none"), bulk.features(long_row, norm="none"))
    wide = _fill(dev, 3, 16000 * 9 + 4, episode=5)
    view = wide[:, : 16000 * 9]                                           # row pitch larger than the row
    assert torch.equal(tma.features(view, norm="none"), bulk.features(view, norm="none"))
    from tal_asrd_b200.streaming import stream_episode
    ep = _fill(dev, 1, 16000 * 95, episode=6)[0]
    a = stream_episode(tma, ep, chunk_seconds=13.0, coalesce_on_device=False)
    b = stream_episode(bulk, ep, chunk_seconds=13.0, coalesce_on_device=False)
    assert torch.equal(a, b)
    y64 = O.logmel_unnormalised_f64(ep[None].cpu().numpy())
    y64 = y64 - y64.mean()
    assert rel_err(a.cpu().numpy(), y64) < TOL


# ------------------------------------------------------------------------------------------ fused normalisation
@pytest.mark.parametrize("shape", [(1, 201), (1, 16000), (3, 24000), (5, 123457), (2, 960000), (64, 480000), (150, 5000)])
def test_fused_normalisation_is_bitwise_the_two_kernel_path(dev, shape):
    """mel -= mel.mean() (tal/asr/models.py:52) inside K1 (cooperative launch, grid barrier, every CTA sweeps its own
    tiles) against K1 followed by sub_scalar_flat_kernel: same reduction order, hence the same bits — also for grids
    smaller than the SM count, ragged last tiles, and repeated launches on one barrier."""
    fused, split = _module(dev, TALFE_FUSED_NORM=2), _module(dev, TALFE_FUSED_NORM=0)     # 2 = fused for every size
    B, L = shape
    x = _fill(dev, B, L, episode=7)
    want = split(x)
    sa, sb = fused.stats_block(dev), split.stats_block(dev)
    for _ in range(3):
        got = fused(x)
        assert torch.equal(got, want)
    fused.features(x, stats=sa)
    split.features(x, stats=sb)
    torch.cuda.synchronize()
    assert torch.equal(sa[:, :3], sb[:, :3])
    raw = fused.features(x, norm="none")
    mu = (sa[0, 1] / sa[0, 0]).float()
    assert torch.equal(raw - mu, want)                                    # exactly "subtract one fp32 scalar"
    assert rel_err(want.cpu().numpy(), O.logmel_f64(x.cpu().numpy())) < TOL if B * L < 4_000_000 else True


def test_fused_normalisation_on_several_streams(dev):
    """Every stream a plan is used on gets its own grid-barrier words; interleaved launches must not disturb each other."""
    mod = _module(dev, TALFE_FUSED_NORM=2)
    x = [_fill(dev, 8, 160000, episode=i) for i in range(4)]
    want = [mod(xi).clone() for xi in x]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    outs = [None] * 4
    for rep in range(3):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                outs[i] = mod(x[i])
        torch.cuda.synchronize()
        for i in range(4):
            assert torch.equal(outs[i], want[i])


# ------------------------------------------------------------------------------------------ drop-in contract
def test_half_module_returns_half_and_feeds_a_half_conv(dev):
    """tal/asr/transcribe.py:255 halves the model and system.py:92 halves the waveform: the features must then be
    float16 (they enter the half TDS conv encoder, models.py:164-167).  A float32 module fed a half waveform returns
    float32, like the reference's type promotion (SURVEY.md §2a [probe])."""
    from tal_asrd_b200 import LogMelSpec, synth
    x = torch.from_numpy(synth.batch(5, 2, 32000)).to(dev)
    m32 = LogMelSpec().to(dev)
    assert m32(x.half()).dtype == torch.float32
    mh = copy.deepcopy(m32).half()
    assert mh.mel_transform.spectrogram.window.dtype == torch.float16       # buffers follow .half() like the reference's
    y = mh(x.half())
    assert y.dtype == torch.float16 and y.shape == (2, 201, 80)
    conv = torch.nn.Conv1d(80, 8, 3).to(dev).half()
    z = conv(y.permute(0, 2, 1))                                            # models.py:167: [B, T, 80] -> [B, 80, T]
    assert z.dtype == torch.float16 and torch.isfinite(z).all()
    # computed in float32 from the UNROUNDED tables, cast once at the end
    want = m32(x.half()).half()
    assert torch.equal(y, want)
    assert mh(x).dtype == torch.float32                                     # float32 waveform into a halved module promotes


def test_module_survives_deepcopy_pickle_and_torch_save_after_forward(dev):
    """The reference module can be copied and pickled at any time (Lightning ddp spawn, EMA replicas); native handles
    therefore live outside the module's state."""
    from tal_asrd_b200 import LogMelSpec, synth
    x = torch.from_numpy(synth.batch(6, 2, 20000)).to(dev)
    m = LogMelSpec().to(dev)
    want = m(x)
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert torch.equal(clone(x), want)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert torch.equal(m2(x), want)
    holder = torch.nn.Module()
    holder.logmelspec = m
    assert torch.equal(copy.deepcopy(holder).logmelspec(x), want)
    del m, m2, holder                                                       # shared plan must outlive every copy
    assert torch.equal(clone(x), want)


def test_under_the_reference_callers_extract_features(dev):
    """ASRModel.extract_features / SDModel.extract_features (tal/asr/models.py:154-162, 430-438), body restated here
    around tal_asrd_b200.LogMelSpec: eval mode = features as they are; training mode = time_mask(freq_mask(x)) with
    Python's global `random`, checked against the masks frozen from the reference's own functions."""
    from tal_asrd_b200 import LogMelSpec, specaug

    class Caller(torch.nn.Module):                                          # the 12 lines of models.py:93 + 154-162
        def __init__(self, n_mels=80):
            super().__init__()
            self.logmelspec = LogMelSpec(n_mels=n_mels)

        def extract_features(self, x, specaug_on=True):
            x = self.logmelspec(x)
            if self.training and specaug_on:
                fb, tb = specaug.sample_masks(x.shape[0], x.shape[1], x.shape[2])     # same `random` draws as the reference
                x = specaug.apply_masks_reference(x, fb, tb)                            # = time_mask(freq_mask(x))
            return x

    z = np.load(os.path.join(GOLDEN_DIR, "specaug_masks.npz"))
    shape = tuple(int(v) for v in z["shape"])                              # (B, T, 80) the masks were frozen for
    n = int(np.prod(shape))
    L = 160 * (shape[1] - 1)
    model = Caller().to(dev)
    x = _fill(dev, shape[0], L, episode=11)
    model.eval()
    plain = model.extract_features(x)
    assert plain.shape == shape and rel_err(plain.cpu().numpy(), O.logmel_f64(x.cpu().numpy())) < TOL
    model.train()
    for key in [k for k in z.files if k.startswith("seed_")][:6]:
        want_zero = np.unpackbits(z[key])[:n].reshape(shape).astype(bool)
        random.seed(int(key.split("_")[1]))
        got = model.extract_features(x)
        assert torch.equal(got == 0, torch.from_numpy(want_zero).to(dev) | (plain == 0))
        keep = ~torch.from_numpy(want_zero).to(dev)
        assert torch.equal(got[keep], plain[keep])
        # the fused form (bands zeroed inside the normalisation sweep) gives the same tensor
        random.seed(int(key.split("_")[1]))
        fb, tb = specaug.sample_masks(*shape)
        fused = model.logmelspec.features(x, spec_augment=(fb, tb))
        assert torch.equal(fused == 0, got == 0) and float((fused - got).abs().max()) < 1e-6
    # encode_features' first op (models.py:167) on the drop-in's output
    assert plain.permute(0, 2, 1).shape == (shape[0], 80, shape[1])


# ------------------------------------------------------------------------------------------ full-size parity vs the oracle
def test_config2_every_row_against_the_float64_oracle(dev):
    """BASELINE configs[1] (64 x 30 s) in full: all 64 rows of the un-normalised features and the scalar mean."""
    from tal_asrd_b200 import LogMelSpec
    mod = LogMelSpec().to(dev)
    x = _fill(dev, 64, 480000, episode=0)
    raw = mod.features(x, norm="none").cpu().numpy()
    y = mod(x).cpu().numpy()
    xs = x.cpu().numpy()
    total = 0.0
    for r in range(64):
        ref = O.logmel_unnormalised_f64(xs[r:r + 1])
        assert rel_err(raw[r:r + 1], ref) < TOL, r
        total += ref.sum()
    mean = total / raw.size
    assert abs(float((raw.astype(np.float64) - y).mean()) - mean) < 1e-5
    assert rel_err(y, raw.astype(np.float64) - mean) < 2e-5


def test_hour_long_episode_streamed_windows_against_the_oracle(dev):
    """BASELINE configs[2]: one hour streamed in chunks (default chunk length and 30 s), 24 random 2-second windows of
    the result — including both reflected ends — against the float64 oracle of exactly those samples; the mean over
    the whole hour against the float64 sum of the device's own un-normalised features."""
    from tal_asrd_b200 import LogMelSpec
    from tal_asrd_b200.streaming import stream_episode
    mod = LogMelSpec().to(dev)
    L = 57_600_000
    T = 1 + L // 160
    ep = _fill(dev, 1, L, episode=42)[0]
    host = ep.cpu().pin_memory()
    raw = stream_episode(mod, host, device=dev, normalise=False)           # from pinned host memory, default chunks
    raw30 = stream_episode(mod, ep, 30.0, normalise=False, coalesce_on_device=False)                 # device resident, 120 chunks
    one = mod.features(ep[None], norm="none")
    # chunking never changes a bit of any frame that has its pair partner; the episode's LAST frame (T is odd) shares its
    # packed rows 18 / 19 with a partner frame that does not exist, whose (discarded) samples differ between the paths:
    # that one frame agrees to the last bit or two (DESIGN.md §3 "known deliberate differences")
    for got in (raw, raw30):
        assert torch.equal(got[:, :T - 1], one[:, :T - 1])
        assert float((got[:, T - 1] - one[:, T - 1]).abs().max()) < 1e-6
    y = stream_episode(mod, host, device=dev)
    mean64 = float(one.double().mean())
    assert float((one - y).double().mean() - mean64) < 1e-6 and float(((one - mean64) - y).abs().max()) < 2e-5
    xs = host.numpy()
    rng = np.random.default_rng(3)
    starts = [0, T - 200] + [int(v) for v in rng.integers(2, T - 204, size=22)]
    got = raw[0].cpu().numpy()
    for t0 in starts:
        t1 = t0 + 200                                                      # 2 s of frames
        if t0 == 0:                                                        # left edge: the oracle's own reflection is the episode's
            seg = xs[: 160 * (t1 + 2)]
            ref = O.logmel_unnormalised_f64(seg[None])[0][:200]
        elif t1 == T:                                                      # right edge: keep the true end of the episode
            s0 = 160 * (t0 - 2)
            ref = O.logmel_unnormalised_f64(xs[None, s0:])[0][2:]
        else:                                                              # interior: two frames of context on both sides
            s0 = 160 * (t0 - 2)
            ref = O.logmel_unnormalised_f64(xs[None, s0: s0 + 160 * 204])[0][2:202]
        assert ref.shape[0] == 200, (t0, ref.shape)
        assert rel_err(got[t0:t1], ref) < TOL, t0


def test_config4_every_row_against_the_oracle(dev):
    """BASELINE configs[3]: ragged batch 1 s .. 10 min; per-row semantics for EVERY row against the float64 oracle of
    that row alone, padded semantics against the oracle on the padded batch rows (tail frames = log eps)."""
    from tal_asrd_b200 import LogMelSpec, _lib
    lib = _lib.load()
    mod = LogMelSpec().to(dev)
    rng = np.random.default_rng(4)
    lens = np.exp(rng.uniform(np.log(16000), np.log(9_600_000), size=12)).astype(np.int64)
    lens[0], lens[-1] = 16000, 9_600_000
    Lmax = int(lens.max())
    x = torch.zeros(len(lens), Lmax, device=dev)
    for r, n in enumerate(lens):
        _lib.check(lib.talfe_synth_fill(x[r].data_ptr(), _lib.F32, 1, int(n), int(n), 2020, 100 + r, 0, None))
    per_row = mod.features(x, audio_lens=torch.from_numpy(lens), norm="row").cpu().numpy()
    packed, offs = mod.features_packed(x, torch.from_numpy(lens), norm="row")
    packed, offs = packed.cpu().numpy(), offs.cpu().numpy()
    padded_raw = mod.features(x, norm="none").cpu().numpy()
    xs = x.cpu().numpy()
    for r, n in enumerate(lens):
        ref, frames = O.logmel_rows_f64([xs[r, :int(n)]], mode="row")
        Tr = frames[0]
        assert Tr == 1 + int(n) // 160
        assert rel_err(per_row[r, :Tr], ref[0]) < TOL, r
        assert not per_row[r, Tr:].any()
        assert rel_err(packed[offs[r]:offs[r + 1]], ref[0]) < TOL, r
        # padded semantics (the reference's): the row as the collater hands it over, zero tail included
        refp = O.logmel_unnormalised_f64(xs[r:r + 1])[0]
        assert rel_err(padded_raw[r], refp) < TOL, r


# ------------------------------------------------------------------------------------------ given statistics (dataset-level CMVN)
@pytest.mark.parametrize("kernel", ["ws", "legacy", "fl"])
@pytest.mark.parametrize("norm", ["row_mel_var", "row_mel", "batch"])
def test_given_statistics_in_the_kernel_equal_the_sweep_bitwise(dev, kernel, norm):
    """talfe_job::given_stats — corpus pass 2: every row normalised with ONE all-reduced statistics block inside the
    transform kernel — against the two-step form (un-normalised transform, then talfe_apply_stats with that block), which
    tests/test_streaming_corpus.py pins to the float64 oracle.  Same two roundings per value, hence the same bits: dense,
    ragged (zero fill beyond a row's own length) and [B, 80, T] outputs, fp32 and int16 PCM."""
    m = _module(dev, TALFE_KERNEL=kernel)
    x = _fill(dev, 5, 16000 * 7 + 123, episode=3)
    # a statistics block from a DIFFERENT batch (the corpus-level sums are not this call's own)
    other = _fill(dev, 3, 16000 * 4, episode=11)
    blocks = m.stats_block(dev, rows=3)
    m.features(other, norm="row_mel_var", stats=blocks, defer_normalise=True)
    block = blocks.sum(dim=0, keepdim=True).contiguous()
    lens = torch.tensor([x.shape[1], 50000, 201, 99999, 16000], device=dev)
    for layout in ("tm", "mt"):
        for audio_lens in (None, lens):
            for xx in (x, (x * 32767).round().to(torch.int16)):
                want = m.features(xx, audio_lens=audio_lens, norm="none", layout=layout)
                if norm == "batch":
                    # apply_stats(batch) uses block 0 for every row as well
                    m.apply_stats(want, block, norm="batch", layout=layout,
                                  valid_frames=None if audio_lens is None else 1 + audio_lens // 160)
                else:
                    rows = block.expand(xx.shape[0], -1).contiguous()
                    m.apply_stats(want, rows, norm=norm, layout=layout,
                                  valid_frames=None if audio_lens is None else 1 + audio_lens // 160)
                got = m.features(xx, audio_lens=audio_lens, norm=norm, layout=layout, given_stats=block)
                assert torch.equal(got, want), (kernel, norm, layout, audio_lens is not None, xx.dtype)
    with pytest.raises(ValueError):
        m.features(x, norm="none", given_stats=block)


def test_given_statistics_generic_geometry(dev):
    """The same through the generic-geometry kernel (LogMelSpec(sr != 16000)): the sweep applies the block."""
    from tal_asrd_b200 import LogMelSpec
    m = LogMelSpec(sr=8000).to(dev)
    x = _fill(dev, 3, 8000 * 5, episode=5)
    blocks = m.stats_block(dev, rows=3)
    m.features(x, norm="row_mel_var", stats=blocks, defer_normalise=True)
    block = blocks.sum(dim=0, keepdim=True).contiguous()
    want = m.features(x, norm="none")
    m.apply_stats(want, block.expand(3, -1).contiguous(), norm="row_mel_var")
    assert torch.equal(m.features(x, norm="row_mel_var", given_stats=block), want)


def test_compact_tile_list_edge_cases(dev):
    """Ragged calls through the compact tile list (ws kernel) against the legacy kernel, which visits the full tile grid:
    rows without a single frame (length <= 200), rows that end inside a tile, one row that fills the whole buffer, a
    batch whose first / last rows are empty, padded and [B, 80, T] outputs, un-normalised (bitwise) and per-row CMVN."""
    ws, legacy = _module(dev, TALFE_KERNEL="ws"), _module(dev, TALFE_KERNEL="legacy")
    x = _fill(dev, 7, 16000 * 6 + 55, episode=31)
    L = x.shape[1]
    for lens in ([100, L, 5120 + 200, 0, 201, 32 * 160 * 3, 150], [L] * 7, [200, 200, 200, 200, 200, 200, 4000]):
        al = torch.tensor(lens, device=dev)
        for layout in ("tm", "mt"):
            a = ws.features(x, audio_lens=al, norm="none", layout=layout)
            b = legacy.features(x, audio_lens=al, norm="none", layout=layout)
            assert torch.isfinite(a).all() and torch.equal(a, b), (lens, layout)
            a = ws.features(x, audio_lens=al, norm="row_mel", layout=layout)
            b = legacy.features(x, audio_lens=al, norm="row_mel", layout=layout)
            assert torch.isfinite(a).all() and float((a - b).abs().max()) < 1e-5, (lens, layout)
    # the output buffer is reused: padding frames of a short row must be re-zeroed by the call itself
    out = torch.full((7, 1 + L // 160, 80), 7.0, device=dev)
    al = torch.tensor([16000, 100, L, 3000, 9000, 201, 48000], device=dev)
    ws.features(x, audio_lens=al, norm="none", out=out)
    want = legacy.features(x, audio_lens=al, norm="none")
    assert torch.equal(out, want)


@pytest.mark.parametrize("kernel", ["ws", "legacy"])
def test_lens_as_padding_hint_is_the_reference_result(dev, kernel):
    """talfe_job::lens_are_padding_hint: a zero-padded ragged batch (what the reference's collaters hand over) with the
    REFERENCE semantics — every row has the padded length's frames, the padding frames are log(eps) and count in the mean —
    computed only where a frame can see a real sample.  Un-normalised: bitwise the plain call on the same buffer (the constant
    fill is the kernel's own expression for an all-zero frame); batch mean: to rounding.  Rows that end just short of the
    padded length are reached by the reflection at the padded end and must be computed in full."""
    m = _module(dev, TALFE_KERNEL=kernel)
    L = 16000 * 9 + 37
    x = _fill(dev, 8, L, episode=41)
    lens = [L, 16000, 0, L - 1, L - 150, L - 450, 5000, 777]
    for r, n in enumerate(lens):
        x[r, n:] = 0.0
    al = torch.tensor(lens, device=dev)
    for xx in (x, (x * 32767).round().to(torch.int16)):
        for layout in ("tm", "mt"):
            want = m.features(xx, norm="none", layout=layout)
            got = m.features(xx, audio_lens=al, lens_are_padding=True, norm="none", layout=layout)
            assert torch.equal(got, want), (layout, xx.dtype)
            want = m.features(xx, norm="batch", layout=layout)
            got = m.features(xx, audio_lens=al, lens_are_padding=True, norm="batch", layout=layout)
            assert float((got - want).abs().max()) < 2e-6, (layout, xx.dtype)
    # and against the float64 oracle of the reference semantics
    ref = O.logmel_f64(x.cpu().numpy())
    got = m.features(x, audio_lens=al, lens_are_padding=True).cpu().numpy()
    assert rel_err(got, ref) < TOL
    with pytest.raises(ValueError):
        m.features(x, audio_lens=al, lens_are_padding=True, norm="row")


def test_drop_in_with_detected_padding(dev):
    """LogMelSpec(detect_padding=True): forward(audio) finds the collater's zero padding itself (talfe_detect_padding) and
    computes only what can see a real sample — the reference's result with the reference's call."""
    from tal_asrd_b200 import LogMelSpec
    plain, fast = LogMelSpec().to(dev), LogMelSpec(detect_padding=True).to(dev)
    L = 16000 * 8 + 3
    x = _fill(dev, 6, L, episode=51)
    lens = [L, 12345, 0, L - 100, 64000, 1]
    for r, n in enumerate(lens):
        x[r, n:] = 0.0
    x[4, 100:200] = 0.0                                                  # zeros inside a row are not padding
    for xx in (x, x.half(), (x * 32767).round().to(torch.int16)):
        got_lens = fast.padding_lens(xx).cpu().tolist()
        want_lens = [int(torch.nonzero(row).max()) + 1 if bool((row != 0).any()) else 0 for row in xx]
        assert got_lens == want_lens
        assert float((fast(xx) - plain(xx)).abs().max()) < 2e-6
    dense = _fill(dev, 4, 48000, episode=52)
    assert float((fast(dense) - plain(dense)).abs().max()) < 2e-6
    assert copy.deepcopy(fast).detect_padding
