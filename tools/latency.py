"""Latency of small calls (BASELINE config 1: one 60 s clip) and batch-size sweep on one B200."""
import sys, json, torch
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, _lib
dev = torch.device('cuda:0'); lib = _lib.load(); mod = LogMelSpec().to(dev)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
res = {}
for B, secs in ((1, 60), (1, 1), (1, 10), (8, 30), (16, 30), (32, 30), (64, 30), (128, 30), (256, 30), (64, 10), (64, 20)):
    L = secs * 16000
    w = torch.empty(B, L, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, B, L, L, 2020, 0, 0, None))
    out = torch.empty(B, 1 + L // 160, 80, device=dev)
    us = timeit(lambda: mod.features(w, out=out))
    fr = B * (1 + L // 160)
    res[f"{B}x{secs}s"] = {"us_per_call": round(us, 2), "frames": fr, "Mframes_per_s": round(fr / us, 1)}
print(json.dumps(res, indent=1))
