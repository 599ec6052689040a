"""CPU: bank-conflict audit of the warp-specialised kernel's shared-memory layouts (csrc/talfe_core.cuh,
"Warp-specialised path").  The addresses come from the emulator library, i.e. from the same expressions
the kernel's stage functions use; the model is the hardware's: 32 banks of 4 bytes, a 32-bit access is
served per warp, a 64-bit access per half-warp, a 128-bit access per quarter-warp, and the number of
wavefronts of a request is the largest number of DISTINCT addresses that fall into one bank."""
import ctypes

import pytest

from test_host_emul import emul  # noqa: F401  (fixture: builds tests/_build/libtalfe_emul.so)

ROLE_THREADS = 320


def wavefronts(addrs, width_words):
    """addrs: per-lane addresses in units of the access width.  Returns wavefronts for one warp instruction."""
    lanes_per_pass = 32 // width_words
    total = 0
    for p0 in range(0, len(addrs), lanes_per_pass):
        group = addrs[p0:p0 + lanes_per_pass]
        banks = {}
        for a in set(group):
            for wd in range(width_words):
                banks.setdefault((a * width_words + wd) % 32, set()).add(a)
        total += max(len(v) for v in banks.values())
    return total


def audit(lib, kind, width_words, idx_range, idx2_range=(0,)):
    lib.talfe_emul_ws_addr.restype = ctypes.c_int64
    lib.talfe_emul_ws_addr.argtypes = [ctypes.c_int] * 4
    worst, total, ideal = 0, 0, 0
    for idx in idx_range:
        for idx2 in idx2_range:
            for w in range(ROLE_THREADS // 32):
                addrs = [lib.talfe_emul_ws_addr(kind, 32 * w + lane, idx, idx2) for lane in range(32)]
                n = wavefronts(addrs, width_words)
                total += n
                ideal += width_words
                worst = max(worst, n / width_words)
    return worst, total, ideal


def test_waveform_loads_conflict_free(emul):
    assert audit(emul, 0, 1, range(28))[0] == 1


def test_exchange_stores_and_loads_conflict_free(emul):
    assert audit(emul, 1, 2, range(20))[0] == 1          # producers, group-major STS.64
    assert audit(emul, 2, 4, range(10))[0] == 1          # consumers, pair-minor LDS.128


def test_power_stores_and_mel_loads_conflict_free(emul):
    lib = emul
    lib.talfe_emul_ws_addr.restype = ctypes.c_int64
    lib.talfe_emul_ws_addr.argtypes = [ctypes.c_int] * 4
    # normal rows: consumer warps 0..8 (threads 0..287), packed rows: warp 9
    for idx in range(20):
        for w in range(9):
            assert wavefronts([lib.talfe_emul_ws_addr(3, 32 * w + lane, idx, 0) for lane in range(32)], 1) == 1
    for idx in range(10):
        assert wavefronts([lib.talfe_emul_ws_addr(4, 288 + lane, idx, 0) for lane in range(32)], 2) == 2
    for slot, width in enumerate((2, 4, 7, 13)):
        assert audit(lib, 5, 2, [slot], range(width))[0] == 1


def test_feature_staging_stores_two_way(emul):
    worst, total, ideal = audit(emul, 6, 1, range(4), range(2))
    assert worst <= 2
