"""Development check of the frame-per-lane kernel (TALFE_KERNEL=fl) against the default kernel on the same inputs.
usage: python tools/fl_debug.py [B] [L]   (run under compute-sanitizer for fault localisation)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib, frontend  # noqa: E402
from tal_asrd_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
dev = torch.device("cuda:0")
x = torch.from_numpy(synth.batch(2020, B, L)).to(dev)
outs = {}
for kern in ("ws", "fl"):
    os.environ["TALFE_KERNEL"] = kern
    _lib._LIB = None
    frontend._PLANS.clear()
    mod = LogMelSpec().to(dev)
    for norm in ("none", "batch"):
        y = mod.features(x, norm=norm)
        torch.cuda.synchronize()
        outs[(kern, norm)] = y.clone()
        print(kern, norm, tuple(y.shape), float(y.abs().max()), flush=True)
for norm in ("none", "batch"):
    a, b = outs[("ws", norm)], outs[("fl", norm)]
    d = (a - b).abs()
    print(norm, "max abs diff", float(d.max()), "at", [int(v) for v in torch.nonzero(d == d.max())[0]], "mean", float(d.mean()), flush=True)
    bad = torch.nonzero(d > 1e-4)
    print("  entries > 1e-4:", int(bad.shape[0]), bad[:8].tolist())
