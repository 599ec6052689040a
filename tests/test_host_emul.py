"""CPU: drive the kernel's own __host__ __device__ stage functions (csrc/talfe_core.cuh) through the
host emulator (csrc/host_emul.cu) and compare with the oracle.  This checks the prime-factor FFT-20
index maps, the real-structured 20 x 20 split, table layouts and the mel/log stage without a GPU.
The emulator is test-only; the product library contains none of it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case, rel_err
from oracle import logmel_oracle as O

BUILD = os.path.join(ROOT, "tests", "_build")
SRC = os.path.join(ROOT, "tal_asrd_b200", "csrc", "host_emul.cu")
LIB = os.path.join(BUILD, "libtalfe_emul.so")


@pytest.fixture(scope="module")
def emul():
    from tal_asrd_b200 import _build
    os.makedirs(BUILD, exist_ok=True)
    deps = [SRC, os.path.join(os.path.dirname(SRC), "talfe_core.cuh"), os.path.join(os.path.dirname(SRC), "talfe_tables.h")]
    if not os.path.isfile(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([_build.nvcc_path(), "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                        "-gencode", "arch=compute_100a,code=sm_100a", SRC, "-o", LIB], check=True)
    lib = ctypes.CDLL(LIB)
    lib.talfe_emul_logmel.restype = ctypes.c_int
    lib.talfe_emul_logmel.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    return lib


def emul_logmel(lib, x, n_mels=80, tables=True, generic=0):
    x = np.ascontiguousarray(x, np.float32)
    T = 1 + x.shape[0] // 160
    out = np.zeros((T, n_mels), np.float32)
    if tables:
        win, fb = O.reference_tables(n_mels)
        rc = lib.talfe_emul_logmel(x.ctypes.data, x.shape[0], n_mels, win.ctypes.data, fb.ctypes.data, 1e-6, generic, out.ctypes.data)
    else:
        rc = lib.talfe_emul_logmel(x.ctypes.data, x.shape[0], n_mels, None, None, 1e-6, generic, out.ctypes.data)
    assert rc == 0
    return out


def test_fft20_prime_factor_maps(emul):
    rng = np.random.default_rng(1)
    for _ in range(5):
        z = rng.standard_normal(20) + 1j * rng.standard_normal(20)
        buf = np.empty(40, np.float32)
        buf[0::2], buf[1::2] = z.real, z.imag
        emul.talfe_emul_fft20(buf.ctypes.data_as(ctypes.c_void_p))
        got = buf[0::2] + 1j * buf[1::2]
        assert np.abs(got - np.fft.fft(z)).max() < 5e-6
    # unit impulses pin every output index individually
    for n in range(20):
        buf = np.zeros(40, np.float32)
        buf[2 * n] = 1.0
        emul.talfe_emul_fft20(buf.ctypes.data_as(ctypes.c_void_p))
        want = np.exp(-2j * np.pi * n * np.arange(20) / 20)
        assert np.abs((buf[0::2] + 1j * buf[1::2]) - want).max() < 1e-6


@pytest.mark.parametrize("name", ["lcg_noise", "tone_1k", "dc_half", "len_201", "len_400", "len_15999", "len_16001",
                                  "loud_fullscale", "tiny_amplitude", "zeros"])
def test_emulated_group_matches_reference_double(emul, name):
    c = load_case(name)
    got = emul_logmel(emul, c["audio"][0])
    ref = c["ref_f64_unnormalised"][0]
    gap = rel_err(c["ref_f32"], c["ref_f64"])
    assert rel_err(got, ref) <= max(1e-4, 2 * gap)


def test_emulated_other_mel_counts_and_builtin_tables(emul):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(8000) * 0.1).astype(np.float32)
    for n_mels in (80, 64, 40, 23):
        got = emul_logmel(emul, x, n_mels)
        ref = O.logmel_unnormalised_f64(x[None], n_mels=n_mels)[0]
        assert rel_err(got, ref) < 1e-4, n_mels
    # the runtime-width mel loop (used for non-reference layouts) agrees with the unrolled one on 80 mels
    assert np.array_equal(emul_logmel(emul, x, 80, generic=1), emul_logmel(emul, x, 80))
    # tables computed inside the C library (no torch): within tolerance of the reference's fp32 tables
    assert rel_err(emul_logmel(emul, x, 80, tables=False), O.logmel_unnormalised_f64(x[None])[0]) < 1e-4


# ---- warp-specialised kernel (csrc/talfe_ws.cuh): same stage functions and E / P / Y layouts, a tile at a time
def emul_logmel_ws(lib, x):
    lib.talfe_emul_logmel_ws.restype = ctypes.c_int
    lib.talfe_emul_logmel_ws.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                         ctypes.c_void_p]
    x = np.ascontiguousarray(x, np.float32)
    T = 1 + x.shape[0] // 160
    out = np.zeros((T, 80), np.float32)
    win, fb = O.reference_tables(80)
    rc = lib.talfe_emul_logmel_ws(x.ctypes.data, x.shape[0], win.ctypes.data, fb.ctypes.data, 1e-6, out.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("name", ["lcg_noise", "tone_1k", "dc_half", "len_201", "len_15999", "loud_fullscale", "tiny_amplitude",
                                  "zeros"])
def test_emulated_ws_tile_matches_reference_double(emul, name):
    c = load_case(name)
    got = emul_logmel_ws(emul, c["audio"][0])
    ref = c["ref_f64_unnormalised"][0]
    gap = rel_err(c["ref_f32"], c["ref_f64"])
    assert np.isfinite(got).all()
    assert rel_err(got, ref) <= max(1e-4, 2 * gap)


def test_emulated_ws_equals_legacy_layout_bitwise(emul):
    """Both kernels run the same arithmetic in the same order; only the shared-memory layouts differ."""
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(16000 * 3 + 77) * 0.1).astype(np.float32)
    a = emul_logmel(emul, x)
    b = emul_logmel_ws(emul, x)
    assert np.array_equal(a, b)
