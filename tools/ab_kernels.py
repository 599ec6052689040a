"""A/B timing of the K1 variants on one GPU (64 x 30 s, inputs resident in HBM, 3 rotating batches).
Usage: python tools/ab_kernels.py [steps]  ->  one JSON object on stdout."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib  # noqa: E402

B, L = 64, 480000
T = 1 + L // 160
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
lib = _lib.load()
waves = []
for i in range(3):
    w = torch.empty(B, L, dtype=torch.float32, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, B, L, L, 2020, i * B, 0, torch.cuda.current_stream().cuda_stream))
    waves.append(w)
outs = [torch.empty(B, T, 80, dtype=torch.float32, device=dev) for _ in range(2)]


def make(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        m = LogMelSpec().to(dev)
        m.plan(dev)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return m


def timeit(mod, norm):
    for i in range(5):
        mod.features(waves[i % 3], norm=norm, out=outs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        mod.features(waves[i % 3], norm=norm, out=outs[i % 2])
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, host_ms


variants = {
    "legacy": {"TALFE_KERNEL": "legacy"},
    "ws": {"TALFE_KERNEL": "ws"},
    "ws_l2pf": {"TALFE_KERNEL": "ws", "TALFE_L2_PREFETCH": "1"},
}
res = {}
ref = None
for name, env in variants.items():
    mod = make(env)
    y = mod.features(waves[0], norm="none").clone()
    torch.cuda.synchronize()
    if ref is None:
        ref = y
    k_ms, k_host = timeit(mod, "none")
    f_ms, f_host = timeit(mod, "batch")
    res[name] = {"kernel_ms": k_ms, "forward_ms": f_ms, "host_ms_per_call": k_host,
                 "gframes_per_s_kernel": B * T / k_ms / 1e6, "equal_to_legacy": bool(torch.equal(y, ref)), "max_abs_diff_vs_legacy": float((y - ref).abs().max()),
                 "frac_of_6445GBs": B * T * 960 / (k_ms * 1e-3) / 6445.3e9}
print(json.dumps(res, indent=1))
