"""CPU model checks of the frame-per-lane kernel's layouts (tal_asrd_b200/csrc/talfe_fl.cuh): the constants are read
from the header, the access patterns are the ones the kernel issues.  No GPU needed."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tal_asrd_b200", "csrc", "talfe_fl.cuh")
SRC = open(HDR).read()
HOP, NFFT = 160, 400


def const(name):
    m = re.search(r"constexpr int %s = ([^;]+);" % name, SRC)
    assert m, name
    expr = m.group(1).split("//")[0]
    expr = expr.replace("kHop", str(HOP)).replace("kNfft", str(NFFT)).replace("kMaxMels", "80").replace("(int)sizeof(float)", "4").replace("(int)sizeof(double2)", "16")
    for other in re.findall(r"kFl[A-Za-z0-9]+", expr):
        expr = expr.replace(other, str(const(other)))
    return int(eval(expr.replace("/", "//")))


def groups_128bit(byte_addrs):
    """16-byte bank groups (8 of them) touched by one quarter-warp phase of a 128-bit access."""
    return [(a // 16) % 8 for a in byte_addrs]


def test_waveform_tile_reads_are_conflict_free():
    pitch = const("kFlRowPitch")
    assert pitch == 164 and (pitch * 4) % 16 == 0          # rows the copy engine can write, lanes 16-byte aligned
    for m in range(20):
        for c4 in range(5):
            i = 20 * m + 4 * c4                             # sample index inside the frame
            word = i + 4 * (i // HOP)
            for q in range(4):                              # the four quarter-warp phases of an LDS.128
                addrs = [4 * (pitch * lane + word) for lane in range(8 * q, 8 * q + 8)]
                assert all(a % 16 == 0 for a in addrs)
                assert len(set(groups_128bit(addrs))) == 8, (m, c4, q)
    # every sample a frame needs is inside the box the tensor copy delivers
    rows, tile = const("kFlRows"), const("kFlTileSamples")
    assert rows * HOP >= tile == HOP * 32 + (NFFT - HOP)
    assert const("kFlSpan") == (rows - 1) * HOP + pitch


def test_feature_staging_is_conflict_free_and_storable():
    pitch = const("kFlYPitch")
    assert pitch == 84 and (pitch * 4) % 16 == 0            # box rows of the tensor store: multiples of 16 bytes
    for m0 in range(0, 80, 4):
        for q in range(4):
            addrs = [4 * (pitch * lane + m0) for lane in range(8 * q, 8 * q + 8)]
            assert len(set(groups_128bit(addrs))) == 8


def test_tensor_memory_columns():
    cols = const("kFlTmemCols")
    row0 = const("kFlColRow0")
    row_col = lambda k1: cols - 40 * k1
    spans = {0: (row0, row0 + 20)}
    spans.update({k1: (row_col(k1), row_col(k1) + 40) for k1 in range(1, 11)})
    used = np.zeros(cols, int)
    for lo, hi in spans.values():
        assert 0 <= lo and hi <= cols
        used[lo:hi] += 1
    assert used.max() == 1                                   # the parked rows do not overlap each other
    p_lo, p_hi = 1, 200 + 16                                 # P[k] at column k, read padding of the widest class
    overlapping = sorted(k1 for k1, (lo, hi) in spans.items() if lo < p_hi and hi > p_lo)
    assert overlapping == [0, 8, 9, 10]                      # exactly the rows stage 2 loads before its first power store
    # those four rows are all in the half that loads them up front (fl_stage2: half 1)
    body = SRC[SRC.index("__device__ __forceinline__ void fl_stage2("):]
    body = body[:body.index("pair_sync();")]                 # everything loaded BEFORE the barrier that precedes the first power store
    first_half, second_half = body.split("} else {", 1)
    assert "fl_row_col(7)" in first_half and "fl_row_col(6)" in first_half
    for k1 in (10, 9, 8):
        assert f"fl_row_col({k1})" in second_half
    assert "kFlColRow0" in second_half and "tmem_wait_ld();" in second_half


def test_shared_memory_budget():
    assert const("kFlSmemBytes") <= 232448
    assert const("kFlXBytes") % 128 == 0 and const("kFlWarpBytes") % 128 == 0   # tensor-copy destinations
