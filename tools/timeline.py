"""Per-warp phase timeline of logmel_ws_kernel on CTA 0 (development build -DTALFE_TIMELINE, tools/build_variant.sh):
average cycles each role spends in each phase of a steady-state tile, and when the phases of the roles happen relative
to each other.   usage: TALFE_LIB=tools/_abl/libtalfe_timeline.so TALFE_ABI_CHECK=0 python tools/timeline.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
mod = LogMelSpec().to(dev)
B, L = 64, 480000
x = torch.empty(B, L, device=dev)
_lib.check(lib.talfe_synth_fill(x.data_ptr(), _lib.F32, B, L, L, 2020, 0, 0, None))
out = torch.empty(B, 1 + L // 160, 80, device=dev)
for _ in range(3):
    mod.features(x, norm="none", out=out)
torch.cuda.synchronize()
buf = np.zeros((20, 64, 8), np.uint32)
lib.talfe_debug_timeline.argtypes = [ctypes.c_void_p]
assert lib.talfe_debug_timeline(buf.ctypes.data) == 0
t = buf.astype(np.int64)
res = {}
k = np.arange(6, 36)                                    # steady-state tiles of CTA 0 (41 in total)
prod = t[:10][:, k, :5]                                 # [warp, tile, slot]
d = np.diff(prod, axis=2) & 0xFFFFFFFF
names = ["wait x_full", "loads + FFT", "wait e_empty", "twiddle + E store"]
res["producer cycles per tile (mean over 10 warps)"] = {n: float(d[:, :, i].mean()) for i, n in enumerate(names)}
res["producer tile period"] = float((np.diff(t[:10][:, k, 0], axis=1) & 0xFFFFFFFF).mean())
res["producer per-warp wait x_full"] = [float(d[w, :, 0].mean()) for w in range(10)]
for grp in (0, 1):
    kk = k[k % 2 == grp]
    cons = t[10 + 5 * grp:15 + 5 * grp][:, kk, :7]
    dc = np.diff(cons, axis=2) & 0xFFFFFFFF
    cn = ["barrier A1", "store issue + wait e_full", "row 0 (LDS, FFT, power, STS)", "row 1", "wait store reads + barrier A2", "mel x 2 + fence"]
    res[f"consumer group {grp} cycles per tile (mean over 5 warps)"] = {n: float(dc[:, :, i].mean()) for i, n in enumerate(cn)}
    res[f"consumer group {grp} tile period"] = float((np.diff(cons[:, :, 0], axis=1) & 0xFFFFFFFF).mean())
    res[f"consumer group {grp} per-warp A1 wait"] = [float(dc[w, :, 0].mean()) for w in range(5)]
    res[f"consumer group {grp} per-warp A2 wait"] = [float(dc[w, :, 4].mean()) for w in range(5)]
# relative timing: when does consumer group q start row 0 of tile k after the producers finished E(k)?
pe = t[:10][:, :, 4].max(axis=0)                        # last producer warp done with tile k
for grp in (0, 1):
    kk = k[k % 2 == grp]
    cs = t[10 + 5 * grp:15 + 5 * grp][:, kk, 2].min(axis=0)
    res[f"group {grp}: first consumer warp past e_full minus last producer arrive (cycles)"] = float(((cs - pe[kk]) & 0xFFFFFFFF).astype(np.int64).mean())
# loader duty: gap between the end of iteration k and the top of iteration k+1, for the warp that has the duty vs the others
gap = (t[:10][:, 1:, 0] - t[:10][:, :-1, 4]) & 0xFFFFFFFF      # [warp, k] : top of k+1 minus end of k
duty, other = [], []
for kk in k:
    for w in range(10):
        (duty if (kk + 2) % 10 == w else other).append(int(gap[w, kk]))     # load_duty(kk + 2) runs at the top of iteration kk + 1
res["gap end(k) -> top(k+1), warp with loader duty"] = {"mean": float(np.mean(duty)), "min": int(np.min(duty)), "max": int(np.max(duty))}
res["gap end(k) -> top(k+1), other warps"] = {"mean": float(np.mean(other)), "min": int(np.min(other)), "max": int(np.max(other))}
# how far apart are the producer warps? spread of the time they finish a tile
fin = t[:10][:, k, 4]
res["producer warps: spread (max - min) of tile completion, cycles"] = float(((fin.max(axis=0) - fin.min(axis=0)) & 0xFFFFFFFF).mean())
lag = (fin - fin.min(axis=0, keepdims=True)) & 0xFFFFFFFF
res["mean completion lag of the warp that had loader duty for tile k+1"] = float(np.mean([lag[(kk + 1) % 10, i] for i, kk in enumerate(k)]))
np.save("gpurun_out/timeline_raw.npy", buf)
print(json.dumps(res, indent=1))
