"""Layout question for a half-size (8 frame pairs) tile of the ws kernel (DESIGN.md §4, round-2 candidate): with
P[bin][g] holding 8 pairs (64 bytes per bin) a half-warp of the mel stage holds TWO mel lanes, whose 64-byte
reads collide unless their bins differ in parity for every tap.  Per slot (width class) the lanes can be permuted
freely and a filter narrower than its slot may start one bin early (zero weight first), so the question per slot is:
can the 20 first bins be made 10 even + 10 odd?  Prints the answer for the reference filterbank."""
import sys

import numpy as np

sys.path.insert(0, '.')
from tal_asrd_b200 import reference_tables

_, fb = reference_tables(80)
fb = fb.numpy()
lo = [int(np.nonzero(fb[:, m])[0][0]) for m in range(80)]
hi = [int(np.nonzero(fb[:, m])[0][-1]) for m in range(80)]
for i in range(4):
    mels = range(20 * i, 20 * i + 20)
    width = max(hi[m] - lo[m] + 1 for m in mels)
    fixed_even = sum(1 for m in mels if lo[m] % 2 == 0 and hi[m] - lo[m] + 1 == width)
    fixed_odd = sum(1 for m in mels if lo[m] % 2 == 1 and hi[m] - lo[m] + 1 == width)
    free = 20 - fixed_even - fixed_odd                     # narrower than the slot: may shift by one bin (lo >= 1 always)
    ok = fixed_even <= 10 and fixed_odd <= 10
    print(f"slot {i}: width {width:2d}  fixed even {fixed_even:2d}  fixed odd {fixed_odd:2d}  shiftable {free:2d}  -> "
          f"{'10 + 10 split exists' if ok else 'NO conflict-free pairing'}")
