"""K1 alone on one hour-long episode: plain, with per-tile partial sums (ROW_MEAN statistics) and with the per-mel column
sums as well (ROW_MEL_* statistics), normalisation deferred — CUDA events, 20 launches each."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib
dev = torch.device("cuda:0")
L = 57_600_000
lib = _lib.load()
pool = []
for i in range(3):
    w = torch.empty(1, L, dtype=torch.float32, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, 1, L, L, 2020, 7 + i, 0, None)); pool.append(w)
mod = LogMelSpec().to(dev)
out = [torch.empty(1, 1 + L // 160, 80, dtype=torch.float32, device=dev) for _ in range(2)]
blk = mod.stats_block(dev)
def t(fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
print("none            us", round(t(lambda i: mod.features(pool[i % 3], norm="none", out=out[i % 2])), 1))
print("row (deferred)  us", round(t(lambda i: mod.features(pool[i % 3], norm="row", stats=blk, defer_normalise=True, out=out[i % 2])), 1))
print("row_mel_var def us", round(t(lambda i: mod.features(pool[i % 3], norm="row_mel_var", stats=blk, defer_normalise=True, out=out[i % 2])), 1))
print("given           us", round(t(lambda i: mod.features(pool[i % 3], norm="row_mel_var", given_stats=blk, out=out[i % 2])), 1))
