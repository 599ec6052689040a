"""One hour-long episode through the corpus pass-1 call (transform + per-mel statistics, normalisation deferred) and the
pass-2 calls (transform, apply_stats) — run under `ncu --metrics gpu__time_duration.sum` for a per-kernel launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib  # noqa: E402

dev = torch.device("cuda:0")
L = 57_600_000
lib = _lib.load()
w = torch.empty(1, L, dtype=torch.float32, device=dev)
_lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, 1, L, L, 2020, 7, 0, None))
mod = LogMelSpec().to(dev)
out = torch.empty(1, 1 + L // 160, 80, dtype=torch.float32, device=dev)
blk = mod.stats_block(dev)
for _ in range(3):
    mod.features(w, norm="row_mel_var", stats=blk, defer_normalise=True, out=out)
    y = mod.features(w, norm="none", out=out)
    mod.apply_stats(y, blk, norm="row_mel_var")
torch.cuda.synchronize()
print("ok")
