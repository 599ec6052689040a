"""CPU: the C-ABI library loads and exports every symbol include/talfe.h declares; host-side logic
(frame counts, tables, argument validation, synthetic audio) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT


def test_library_exports_every_declared_symbol():
    from tal_asrd_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "talfe.h")).read()
    declared = sorted(set(re.findall(r"\b(talfe_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in talfe.h but not exported"
    assert sorted(_lib.EXPORTED) == declared
    assert lib.talfe_version() == _lib.EXPECTED_VERSION == 105
    assert lib.talfe_job_size() == __import__("ctypes").sizeof(_lib.Job)
    assert lib.talfe_strerror(-2).decode().startswith("waveform too short")


def test_job_struct_matches_header_field_order():
    from tal_asrd_b200 import _lib
    header = open(os.path.join(ROOT, "include", "talfe.h")).read()
    body = header[header.index("typedef struct talfe_job {"):header.index("} talfe_job;")]
    fields = re.findall(r"^\s*(?:const\s+)?[a-z_0-9]+\*?\s+\*?([a-z_0-9]+);", body, flags=re.M)
    assert fields == [f[0] for f in _lib.Job._fields_]


def test_num_frames_host_and_abi():
    from tal_asrd_b200 import _lib, num_frames
    lib = _lib.load()
    fc = np.load(os.path.join(GOLDEN_DIR, "frame_counts.npz"))
    for L, T in zip(fc["lengths"], fc["frames"]):
        assert num_frames(int(L)) == int(T) == lib.talfe_num_frames(int(L))
    for L in fc["raises_runtime_error"]:
        assert lib.talfe_num_frames(int(L)) == _lib.ERR_TOO_SHORT
        with pytest.raises(RuntimeError):
            num_frames(int(L))
    assert num_frames(15999) == 100                         # tal/asr/models.py:91


def test_tables_bit_identical_to_reference_buffers():
    from tal_asrd_b200 import reference_tables
    t = np.load(os.path.join(GOLDEN_DIR, "tables.npz"))
    window, fb = reference_tables(80)
    assert np.array_equal(window.numpy(), t["window"]) and np.array_equal(fb.numpy(), t["fb"])


def test_module_contract_without_gpu():
    from tal_asrd_b200 import LogMelSpec
    m = LogMelSpec()
    assert sorted(m.state_dict().keys()) == ["mel_transform.mel_scale.fb", "mel_transform.spectrogram.window"]
    assert not list(m.parameters())
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        m(torch.zeros(2, 16000))                            # never silently falls back to the CPU
    with pytest.raises(ValueError):
        m(torch.zeros(16000))
    # every sample rate the reference's constructor accepts derives its own geometry (models.py:24-32)
    m8 = LogMelSpec(sr=8000, n_mels=40)
    assert (m8.n_fft, m8.hop) == (200, 80) and tuple(m8.mel_transform.mel_scale.fb.shape) == (101, 40)
    assert tuple(LogMelSpec(sr=22050).mel_transform.spectrogram.window.shape) == (551,)
    with pytest.raises(NotImplementedError):
        LogMelSpec(sr=96000)                                # n_fft 2400: beyond what the kernels stage in shared memory
    with pytest.raises(NotImplementedError):
        LogMelSpec(n_mels=81)
    # a reference checkpoint's buffers load under the same keys
    t = np.load(os.path.join(GOLDEN_DIR, "tables.npz"))
    m.load_state_dict({"mel_transform.spectrogram.window": torch.from_numpy(t["window"]),
                       "mel_transform.mel_scale.fb": torch.from_numpy(t["fb"])}, strict=True)


def test_plan_create_without_device_fails_cleanly():
    from tal_asrd_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _lib.load()
    plan = ctypes.c_void_p()
    rc = lib.talfe_plan_create(ctypes.byref(plan), 0, 80, None, None)
    assert rc == -5 and not plan.value                      # TALFE_ERR_CUDA, no fallback
    assert lib.talfe_plan_create(ctypes.byref(plan), 0, 81, None, None) == -3


def test_synth_is_deterministic_and_shaped():
    from tal_asrd_b200 import synth
    a = synth.waveform(2020, 3, 1000, 5000)
    b = synth.waveform(2020, 3, 0, 6000)[1000:]
    assert np.array_equal(a, b)                             # counter-based: any slicing agrees
    x = synth.waveform(2020, 1, 0, 1600000)
    assert x.dtype == np.float32 and np.abs(x).max() < 1.0
    assert 0.02 < x.std() < 0.12
    assert 0.05 < float((x == 0).mean()) < 0.2              # exact-zero gaps
    assert np.array_equal(x * 32768, np.round(x * 32768))   # int16-quantised
