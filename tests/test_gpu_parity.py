"""GPU parity tests: the sm_100a path, called through the C ABI (ctypes -> libtalfe.so), against
the CPU oracle and the frozen outputs of the real reference class.

Tolerance (BASELINE.json north_star): frame counts / indexing bit-exact; feature values within
1e-4 relative, measured as |a - b| <= 1e-4 * max(1, |b|) because values cross zero after the mean
subtraction.  On tonal input the reference's own fp32 result is further than that from exact
arithmetic, so against the float64 twin the bound is max(1e-4, 2 * err(ref32, ref64)).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, golden_case_names, load_case, rel_err
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def make_module(dev, kernel):
    """kernel: 'ws' (warp-specialised, the default for the reference filterbank), 'legacy', or 'fl' (frame per lane,
    tensor memory as transpose scratch: fp32 waveforms; everything else falls through to 'ws'); the library reads
    TALFE_KERNEL when the plan (device tables) is created, so create it here."""
    from tal_asrd_b200 import LogMelSpec
    old = os.environ.get("TALFE_KERNEL")
    os.environ["TALFE_KERNEL"] = kernel
    try:
        m = LogMelSpec().to(dev)
        m.plan(dev)
    finally:
        if old is None:
            os.environ.pop("TALFE_KERNEL", None)
        else:
            os.environ["TALFE_KERNEL"] = old
    return m


@pytest.fixture(scope="module", params=["ws", "legacy", "fl"])
def mod(dev, request):
    return make_module(dev, request.param)


def run(mod, x, **kw):
    xt = torch.as_tensor(x).to("cuda:0")
    y = mod.features(xt, **kw) if kw else mod(xt)
    torch.cuda.synchronize()
    return y


def test_library_is_the_native_one():
    from tal_asrd_b200 import _lib
    lib = _lib.load()
    assert lib.talfe_version() >= 100
    name = os.path.basename(lib._name)                                   # TALFE_LIB may point at an experiment build
    assert name == "libtalfe.so" or (os.environ.get("TALFE_LIB") and name.startswith("libtalfe_"))


@pytest.mark.parametrize("name", golden_case_names())
def test_golden_cases(mod, name):
    c = load_case(name)
    y = run(mod, c["audio"])
    assert y.dtype == torch.float32 and y.is_contiguous() and not y.requires_grad
    assert tuple(y.shape) == c["ref_f32"].shape                      # frame count bit-exact
    got = y.cpu().numpy()
    assert np.isfinite(got).all()
    ref_gap = rel_err(c["ref_f32"], c["ref_f64"])
    assert rel_err(got, c["ref_f64"]) <= max(TOL, 2 * ref_gap), name
    assert rel_err(got, c["ref_f32"]) <= max(TOL, 3 * ref_gap), name


def test_silence_is_exactly_the_floor(mod):
    y = run(mod, np.zeros((2, 4000), np.float32), norm="none")
    assert torch.all(y == y[0, 0, 0])
    assert abs(float(y[0, 0, 0]) - O.LOG_EPS) < 2e-6
    assert float(run(mod, np.zeros((2, 4000), np.float32)).abs().max()) < 1e-6


def test_frame_counts_and_errors(mod, dev):
    fc = np.load(os.path.join(GOLDEN_DIR, "frame_counts.npz"))
    for L, T in zip(fc["lengths"], fc["frames"]):
        y = mod(torch.zeros(1, int(L), device=dev))
        assert y.shape == (1, int(T), 80)
    for L in fc["raises_runtime_error"]:
        with pytest.raises(RuntimeError):
            mod(torch.zeros(1, int(L), device=dev))
    with pytest.raises(ValueError):
        mod(torch.zeros(16000, device=dev))
    with pytest.raises(RuntimeError):
        mod(torch.zeros(1, 16000))                                   # CPU tensor: no CPU path exists


def test_every_frame_index_matches_oracle(mod):
    """Impulse train: each frame's content depends on exact sample indexing incl. both reflected edges."""
    L = 160 * 37 + 123
    x = np.zeros((1, L), np.float32)
    x[0, ::97] = 0.5
    x[0, 1] = -0.7
    x[0, L - 2] = 0.9
    y = run(mod, x, norm="none").cpu().numpy()
    ref = O.logmel_unnormalised_f64(x)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL


def test_batch_scalar_mean_semantics(mod):
    """models.py:52: ONE mean over the whole [B, T, M] tensor, padding frames included."""
    from tal_asrd_b200 import synth
    x = synth.batch(11, 5, 16000 * 3)
    x[3, 20000:] = 0.0                                               # collater-style zero tail
    y = run(mod, x).cpu().numpy().astype(np.float64)
    assert abs(y.mean()) < 2e-6
    ref = O.logmel_f64(x)
    assert rel_err(y, ref) < TOL
    port = O.logmel_port_f32(x).numpy()
    assert rel_err(y, port) < TOL
    alone = run(mod, x[2:3]).cpu().numpy().astype(np.float64)
    diff = y[2] - alone[0]
    assert np.abs(diff - diff.mean()).max() < 2e-5                  # same row differs by a constant only


def test_moderate_batch_against_port(mod):
    from tal_asrd_b200 import synth
    x = synth.batch(2020, 16, 16000 * 10)
    y = run(mod, x).cpu().numpy()
    assert rel_err(y, O.logmel_port_f32(x).numpy()) < TOL
    assert rel_err(y, O.logmel_f64(x)) < TOL


def test_determinism(mod):
    from tal_asrd_b200 import synth
    x = torch.from_numpy(synth.batch(5, 7, 16000 * 4)).cuda()
    a = mod(x).clone()
    b = mod(x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_fp16_and_int16_inputs(mod):
    from tal_asrd_b200 import synth
    pcm = np.stack([synth.pcm16(3, r, 0, 32000) for r in range(3)])
    x32 = pcm.astype(np.float32) / 32768.0
    ref = O.logmel_f64(x32)
    assert rel_err(run(mod, torch.from_numpy(pcm)).cpu().numpy(), ref) < TOL       # int16 PCM
    x16 = torch.from_numpy(x32).half()
    ref16 = O.logmel_f64(x16.float().numpy())
    assert rel_err(run(mod, x16).cpu().numpy(), ref16) < TOL


def test_non_contiguous_and_strided_rows(mod, dev):
    from tal_asrd_b200 import synth
    big = torch.from_numpy(synth.batch(9, 4, 20000)).to(dev)
    view = big[:, 3:16003]                                           # row stride 20000, misaligned start
    y = mod(view)
    ref = O.logmel_f64(view.cpu().numpy())
    assert rel_err(y.cpu().numpy(), ref) < TOL


def test_per_row_lengths(mod):
    """audio_lens given -> each row as if alone (reference run with B=1 per row), zeros beyond."""
    from tal_asrd_b200 import synth
    lens = [16000, 7777, 12345, 201]
    rows = [synth.waveform(1, i, 0, n) for i, n in enumerate(lens)]
    Lmax = max(lens)
    x = np.zeros((len(lens), Lmax), np.float32)
    for i, r in enumerate(rows):
        x[i, :len(r)] = r
    for mode in ("none", "row", "batch", "row_mel", "row_mel_var"):
        y = run(mod, x, audio_lens=torch.tensor(lens), norm=mode).cpu().numpy()
        ref, frames = O.logmel_rows_f64(rows, mode=mode)
        assert y.shape == ref.shape
        tol = TOL if mode != "row_mel_var" else 5e-4               # division by a small std amplifies
        assert rel_err(y, ref) < tol, mode
        for i, f in enumerate(frames):
            assert not y[i, f:].any()


def test_norm_modes_full_rows(mod):
    from tal_asrd_b200 import synth
    x = synth.batch(4, 3, 48000)
    raw = O.logmel_unnormalised_f64(x)
    for mode in ("none", "batch", "row", "row_mel", "row_mel_var"):
        y = run(mod, x, norm=mode).cpu().numpy()
        assert rel_err(y, O.normalise_f64(raw, mode)) < (TOL if mode != "row_mel_var" else 5e-4), mode


def test_mt_layout_is_exact_transpose(mod):
    from tal_asrd_b200 import synth
    x = synth.batch(8, 2, 30000)
    a = run(mod, x, norm="batch", layout="tm")
    b = run(mod, x, norm="batch", layout="mt")
    assert b.shape == (2, 80, a.shape[1])
    assert torch.equal(a.transpose(1, 2).contiguous(), b)


def test_stats_block_and_deferred_normalisation(mod, dev):
    from tal_asrd_b200 import synth
    x = synth.batch(6, 2, 40000)
    stats = mod.stats_block(dev)
    raw = run(mod, x, norm="batch", stats=stats, defer_normalise=True)
    ref_raw = O.logmel_unnormalised_f64(x)
    assert rel_err(raw.cpu().numpy(), ref_raw) < TOL
    s = stats.cpu().numpy()[0]
    assert s[0] == ref_raw.size
    assert abs(s[1] / s[0] - ref_raw.mean()) < 1e-5
    assert abs(s[2] / s[0] - (ref_raw ** 2).mean()) < 1e-3
    mod.apply_stats(raw, stats, norm="batch")
    torch.cuda.synchronize()
    assert rel_err(raw.cpu().numpy(), ref_raw - ref_raw.mean()) < TOL


def test_synth_fill_matches_numpy(dev):
    from tal_asrd_b200 import _lib, synth
    lib = _lib.load()
    for dtype, code in ((torch.float32, _lib.F32), (torch.int16, _lib.I16), (torch.float16, _lib.F16)):
        buf = torch.empty(3, 50000, dtype=dtype, device=dev)
        _lib.check(lib.talfe_synth_fill(buf.data_ptr(), code, 3, 50000, 50000, 2020, 5, 1234,
                                        torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        for r in range(3):
            pcm = synth.pcm16(2020, 5 + r, 1234, 50000)
            want = pcm if dtype == torch.int16 else (pcm.astype(np.float32) / 32768.0).astype(
                np.float16 if dtype == torch.float16 else np.float32)
            assert np.array_equal(buf[r].cpu().numpy(), want)


def test_raw_c_abi_forward(dev):
    """talfe_logmel_forward exactly as a non-Python host would call it: plain pointers and sizes."""
    from tal_asrd_b200 import _lib, reference_tables, synth
    lib = _lib.load()
    window, fb = reference_tables(80)
    plan = ctypes.c_void_p()
    _lib.check(lib.talfe_plan_create(ctypes.byref(plan), 0, 80, window.data_ptr(), fb.data_ptr()))
    x = synth.batch(77, 2, 16000)
    xd = torch.from_numpy(x).to(dev)
    T = lib.talfe_num_frames(16000)
    assert T == 101 and lib.talfe_num_frames(200) == _lib.ERR_TOO_SHORT
    out = torch.empty(2, T, 80, device=dev)
    ws = torch.empty(lib.talfe_workspace_bytes(plan, 2, T), dtype=torch.uint8, device=dev)
    rc = lib.talfe_logmel_forward(plan, xd.data_ptr(), _lib.F32, 2, 16000, 16000, out.data_ptr(), 1e-6,
                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), O.logmel_f64(x)) < TOL
    # error paths: too-small workspace, too-short input; never UB
    assert lib.talfe_logmel_forward(plan, xd.data_ptr(), _lib.F32, 2, 16000, 16000, out.data_ptr(), 1e-6,
                                    ws.data_ptr(), 16, None) == -4
    assert lib.talfe_logmel_forward(plan, xd.data_ptr(), _lib.F32, 2, 200, 16000, out.data_ptr(), 1e-6,
                                    ws.data_ptr(), ws.numel(), None) == -2
    lib.talfe_plan_destroy(plan)


def test_builtin_tables_close_to_reference_tables(dev):
    """Plan built without caller tables (C-only host) stays within tolerance of the reference's."""
    from tal_asrd_b200 import _lib, synth
    lib = _lib.load()
    plan = ctypes.c_void_p()
    _lib.check(lib.talfe_plan_create(ctypes.byref(plan), 0, 80, None, None))
    x = synth.batch(78, 1, 16000)
    xd = torch.from_numpy(x).to(dev)
    out = torch.empty(1, 101, 80, device=dev)
    ws = torch.empty(lib.talfe_workspace_bytes(plan, 1, 101), dtype=torch.uint8, device=dev)
    assert lib.talfe_logmel_forward(plan, xd.data_ptr(), _lib.F32, 1, 16000, 16000, out.data_ptr(), 1e-6,
                                    ws.data_ptr(), ws.numel(), None) == 0
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), O.logmel_f64(x)) < TOL
    lib.talfe_plan_destroy(plan)


def test_full_size_training_chunk_properties(mod, dev):
    """BASELINE config 2 at full size (64 x 30 s): size-independent properties + sampled rows vs oracle."""
    from tal_asrd_b200 import _lib
    lib = _lib.load()
    x = torch.empty(64, 480000, device=dev)
    _lib.check(lib.talfe_synth_fill(x.data_ptr(), _lib.F32, 64, 480000, 480000, 2020, 0, 0, None))
    y = mod(x)
    torch.cuda.synchronize()
    assert y.shape == (64, 3001, 80) and torch.isfinite(y).all()
    assert abs(float(y.double().mean())) < 2e-6                       # zero scalar mean
    raw = mod.features(x, norm="none")
    mu = raw.double().mean()
    assert float((raw - mu.float() - y).abs().max()) < 2e-5           # only a constant was removed
    # gain linearity: x -> 2x adds log(4) wherever mel >> eps
    raw2 = mod.features(2 * x, norm="none")
    loud = raw > -6
    assert float(((raw2 - raw)[loud] - np.log(4.0)).abs().max()) < 1e-3
    # sampled rows against the float64 oracle
    for r in (0, 31, 63):
        ref = O.logmel_unnormalised_f64(x[r:r + 1].cpu().numpy())
        assert rel_err(raw[r:r + 1].cpu().numpy(), ref) < TOL
    # time-shift by one hop: interior frames move by one index
    xs = x[:2, 160:]
    ys = mod.features(xs.contiguous(), norm="none")
    assert float((ys[:, 2:2000] - raw[:2, 3:2001]).abs().max()) < 2e-4


def test_ragged_batch_config4(mod, dev):
    """BASELINE config 4: variable-length utterances (1 s .. 10 min), both padding semantics."""
    from tal_asrd_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(4)
    lens = np.exp(rng.uniform(np.log(16000), np.log(9_600_000), size=12)).astype(np.int64)
    lens[0], lens[-1] = 16000, 9_600_000
    Lmax = int(lens.max())
    x = torch.zeros(len(lens), Lmax, device=dev)
    for r, n in enumerate(lens):                                         # collater layout: zero right-pad
        _lib.check(lib.talfe_synth_fill(x[r].data_ptr(), _lib.F32, 1, int(n), int(n), 2020, 100 + r, 0, None))
    lens_t = torch.from_numpy(lens)
    # (i) padded semantics == the reference on the padded batch: every row gets 1 + Lmax//160 frames
    padded = mod(x)
    assert padded.shape == (len(lens), 1 + Lmax // 160, 80)
    assert abs(float(padded.double().mean())) < 2e-6
    # (ii) per-row semantics: each row equals the front end run on that row alone
    per_row = mod.features(x, audio_lens=lens_t, norm="row")
    for r in (0, 3, len(lens) - 1):
        n = int(lens[r])
        alone = mod(x[r:r + 1, :n].contiguous())
        T = 1 + n // 160
        assert float((per_row[r, :T] - alone[0]).abs().max()) < 2e-5
        assert not per_row[r, T:].any()
    # short row against the float64 oracle end to end
    r = int(np.argmin(lens))
    ref, frames = O.logmel_rows_f64([x[r, :int(lens[r])].cpu().numpy()], mode="row")
    assert rel_err(per_row[r, :frames[0]].cpu().numpy(), ref[0]) < TOL


def test_invalid_arguments_fail_cleanly(mod, dev):
    x = torch.zeros(2, 16000, device=dev)
    with pytest.raises(ValueError):
        mod.features(x, audio_lens=torch.tensor([16000]))               # one length per row required
    with pytest.raises(KeyError):
        mod.features(x, norm="bogus")
    with pytest.raises(ValueError):
        mod.features(x, out=torch.empty(2, 100, 80, device=dev))          # wrong out shape
    y = mod(torch.zeros(1, 201, device=dev))                              # shortest legal input
    assert y.shape == (1, 2, 80)


def test_ws_and_legacy_kernels_agree_bitwise(dev):
    """Same arithmetic in the same order, different work decomposition and shared-memory layouts: un-normalised
    features must be identical bit for bit, run after run (also a cheap race detector for the mbarrier pipeline)."""
    from tal_asrd_b200 import synth
    ws, legacy = make_module(dev, "ws"), make_module(dev, "legacy")
    x = torch.from_numpy(synth.batch(7, 5, 16000 * 21 + 123)).to(dev)          # 66 tiles per row, ragged last tile
    want = legacy.features(x, norm="none")
    for _ in range(5):
        got = ws.features(x, norm="none")
        assert torch.equal(got, want)
    lens = torch.tensor([x.shape[1], 200000, 7777, 201, 150001])
    a = legacy.features(x, audio_lens=lens, norm="none")
    b = ws.features(x, audio_lens=lens, norm="none")
    assert torch.equal(a, b)
    # normalised: the per-(tile, warp) partial sums group the frames differently, so the means agree to rounding only
    a = legacy.features(x, audio_lens=lens, norm="row")
    b = ws.features(x, audio_lens=lens, norm="row")
    assert float((a - b).abs().max()) < 2e-6
    a, b = legacy.features(x, norm="none", layout="mt"), ws.features(x, norm="none", layout="mt")
    assert torch.equal(a, b)
    for dt in (torch.float16, torch.int16):
        xx = x.half() if dt == torch.float16 else (x * 32767).round().to(torch.int16)
        assert torch.equal(legacy.features(xx, norm="none"), ws.features(xx, norm="none"))
    # batch statistics: per-CTA partial sums differ in grouping, so the mean agrees to rounding, not bitwise
    sa, sb = legacy.stats_block(dev), ws.stats_block(dev)
    legacy.features(x, stats=sa); ws.features(x, stats=sb)
    torch.cuda.synchronize()
    # count exact; the sums are fp32 per thread and tile before they are widened, so they agree to ~1e-7 relative
    assert sa[0, 0] == sb[0, 0]
    assert torch.allclose(sa[:, 1:3], sb[:, 1:3], rtol=2e-6, atol=0)


def test_ws_kernel_many_tiles_per_cta(dev):
    """Config-2-sized batch (64 x 30 s): every persistent CTA runs ~41 tiles through both buffers of every queue."""
    from tal_asrd_b200 import synth
    ws, legacy = make_module(dev, "ws"), make_module(dev, "legacy")
    x = torch.from_numpy(synth.batch(3, 64, 480000)).to(dev)
    a = legacy.features(x, norm="none")
    for _ in range(3):
        assert torch.equal(ws.features(x, norm="none"), a)
    ya, yb = legacy(x), ws(x)
    assert float((ya - yb).abs().max()) < 1e-5
