"""Kernel-only timing of K1 (64 x 30 s, norm none) for a given build of the library (development experiments:
ablation builds under tools/_abl/).  usage: python tools/time_lib.py <path-to-libtalfe.so> [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import _build  # noqa: E402

_build.LIB_PATH = os.path.abspath(sys.argv[1])
from tal_asrd_b200 import LogMelSpec, _lib  # noqa: E402

steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
B, L = 64, 480000
T = 1 + L // 160
dev = torch.device("cuda:0")
lib = _lib.load()
waves = []
for i in range(3):
    w = torch.empty(B, L, dtype=torch.float32, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, B, L, L, 2020, i * B, 0, torch.cuda.current_stream().cuda_stream))
    waves.append(w)
outs = [torch.empty(B, T, 80, dtype=torch.float32, device=dev) for _ in range(2)]
mod = LogMelSpec().to(dev)
for i in range(5):
    mod.features(waves[i % 3], norm="none", out=outs[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    mod.features(waves[i % 3], norm="none", out=outs[i % 2])
e1.record()
torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(sys.argv[1]), "kernel_us": 1e3 * e0.elapsed_time(e1) / steps}))
