"""Offline search for the shared-memory layout of the power array that minimises bank conflicts of the
mel-stage reads (LDS.64 of (P_a, P_b)[bin]) while keeping the stage-2 writes conflict-free.
Model: 32 banks x 4 B; a 64-bit access is served per half-warp; wavefronts = max number of distinct
addresses falling in the same 8-byte bank."""
import itertools, sys
import numpy as np
sys.path.insert(0, '.')
from tal_asrd_b200 import reference_tables

_, fb = reference_tables(80)
fb = fb.numpy()
lo = [int(np.nonzero(fb[:, m])[0][0]) for m in range(80)]
hi = [int(np.nonzero(fb[:, m])[0][-1]) for m in range(80)]
R = [max(hi[m] - lo[m] + 1 for m in range(20 * i, 20 * i + 20)) for i in range(4)]
print("widths", R)

def read_cost(pi, gshift, owner=lambda j, i: j + 20 * i):
    """total wavefronts of the mel reads per 80 threads (5 half-warps pattern), all slots / r."""
    total = ideal = 0
    for i in range(4):
        for r in range(R[i]):
            for hw in range(5):
                lanes = range(16 * hw, 16 * hw + 16)
                banks = {}
                for t in lanes:
                    g, j = divmod(t, 20)
                    k = lo[owner(j, i)] + r
                    addr = pi[k] + gshift * g            # in 8-byte units; different groups never share addresses
                    banks.setdefault(addr % 16, set()).add((g, addr))
                total += max(len(v) for v in banks.values())
                ideal += 1
    return total, ideal

def make_pi(sigma):
    pi = {}
    for k in range(0, 230):
        b = min(k // 10, len(sigma) - 1)
        pi[k] = k + sigma[b]
    return pi

base = make_pi([0] * 23)
for gs in (9, 4, 1, 3, 5, 7, 11, 13, 15):
    print("gshift", gs, read_cost(base, gs))

# coordinate descent over per-block cumulative skews
import random
random.seed(0)
best = None
for gs in (9,):
    for trial in range(30):
        inc = [random.randint(0, 2) for _ in range(23)]
        def cost(inc):
            sig = list(itertools.accumulate(inc))
            return read_cost(make_pi(sig), gs)[0] + 0.3 * sig[-1]
        c = cost(inc)
        improved = True
        while improved:
            improved = False
            for b in range(23):
                for v in range(0, 6):
                    if v == inc[b]: continue
                    trial_inc = inc[:b] + [v] + inc[b + 1:]
                    ct = cost(trial_inc)
                    if ct < c - 1e-9:
                        inc, c, improved = trial_inc, ct, True
        sig = list(itertools.accumulate(inc))
        rc = read_cost(make_pi(sig), gs)
        if best is None or rc[0] + 0.3 * sig[-1] < best[0]:
            best = (rc[0] + 0.3 * sig[-1], rc, gs, sig)
            print("best", best)

# ---- lane permutation within each slot (no skew): thread j of slot i owns mel perms[i][j]
print("---- lane permutation search")
def read_cost_perm(perms, gshift, pi=base):
    total = 0
    for i in range(4):
        for r in range(R[i]):
            for hw in range(5):
                banks = {}
                for t in range(16 * hw, 16 * hw + 16):
                    g, j = divmod(t, 20)
                    k = lo[perms[i][j]] + r
                    addr = pi[k] + gshift * g
                    banks.setdefault(addr % 16, set()).add((g, addr))
                total += max(len(v) for v in banks.values())
    return total

random.seed(1)
ident = [list(range(20 * i, 20 * i + 20)) for i in range(4)]
bestp = None
for gs in (9, 1, 3, 5, 7, 11, 13, 15):
    for trial in range(6):
        perms = [p[:] for p in ident]
        for p in perms: random.shuffle(p)
        c = read_cost_perm(perms, gs)
        improved = True
        while improved:
            improved = False
            for i in range(4):
                for a in range(20):
                    for b in range(a + 1, 20):
                        perms[i][a], perms[i][b] = perms[i][b], perms[i][a]
                        ct = read_cost_perm(perms, gs)
                        if ct < c: c, improved = ct, True
                        else: perms[i][a], perms[i][b] = perms[i][b], perms[i][a]
        if bestp is None or c < bestp[0]:
            bestp = (c, gs, [p[:] for p in perms]); print("best perm", c, "gshift", gs)
print(bestp)
