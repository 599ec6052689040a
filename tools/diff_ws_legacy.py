"""For each build of the library: run the ws and the legacy kernel on the same input (norm none) and describe where
their outputs differ (count, magnitude, which frame parity / mel).  usage: python tools/diff_ws_legacy.py lib1.so ..."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import _build, _lib, synth  # noqa: E402
from tal_asrd_b200 import LogMelSpec  # noqa: E402

dev = torch.device("cuda:0")
x = torch.from_numpy(synth.batch(7, 5, 16000 * 21 + 123)).to(dev)
res = {}
for path in sys.argv[1:]:
    _lib._LIB = None
    _build.LIB_PATH = os.path.abspath(path)
    mods = {}
    for kind in ("ws", "legacy"):
        os.environ["TALFE_KERNEL"] = kind
        m = LogMelSpec().to(dev)
        m.plan(dev)
        mods[kind] = m
    os.environ.pop("TALFE_KERNEL", None)
    a = mods["ws"].features(x, norm="none").cpu().numpy()
    b = mods["legacy"].features(x, norm="none").cpu().numpy()
    d = a != b
    r = {"n_diff": int(d.sum()), "n": int(d.size), "max_abs": float(np.abs(a - b).max())}
    if d.any():
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
        r["max_ulp"] = int(ulp[d].max())
        r["by_frame_parity"] = [int(d[:, 0::2].sum()), int(d[:, 1::2].sum())]
        r["by_mel_top"] = sorted(((int(c), int(m)) for m, c in enumerate(d.sum(axis=(0, 1))) if c), reverse=True)[:12]
        r["mels_with_diff"] = int((d.sum(axis=(0, 1)) > 0).sum())
        idx = np.argwhere(d)[:5]
        r["examples"] = [(list(map(int, i)), float(a[tuple(i)]), float(b[tuple(i)])) for i in idx]
    res[os.path.basename(path)] = r
print(json.dumps(res, indent=1))
