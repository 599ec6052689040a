"""Fused (in-kernel) vs two-launch normalisation across call sizes: device time from a CUDA graph and back-to-back time."""
import json, os, sys, time, torch
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, _lib, frontend
dev = torch.device('cuda:0')
res = {}
for fused in ("0", "2", "0", "2"):
    os.environ["TALFE_FUSED_NORM"] = fused
    _lib._LIB = None; frontend._PLANS.clear()
    lib = _lib.load(); mod = LogMelSpec().to(dev)
    for B, secs in ((1, 1), (1, 30), (1, 60), (4, 30), (8, 30), (16, 30), (32, 30), (64, 30)):
        L = secs * 16000
        w = torch.empty(B, L, device=dev)
        _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, B, L, L, 2020, 0, 0, None))
        out = torch.empty(B, 1 + L // 160, 80, device=dev)
        fn = lambda: mod.features(w, out=out)
        for _ in range(50): fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(200): fn()
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t1) / 200 * 1e6
        g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn(); side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        res.setdefault(f"{B}x{secs}s", {}).setdefault("fused" if fused == "2" else "two launches", []).append(
            {"back_to_back_us": round(t_all, 2), "graph_us": round(e0.elapsed_time(e1) / 200 * 1e3, 2)})
print(json.dumps(res, indent=1))
