"""Where a small call's time goes (BASELINE configs[0]: one 60 s clip): host-side enqueue cost vs device time, with and
without the in-kernel normalisation, back to back and replayed from a CUDA graph."""
import json, os, sys, time, torch
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, _lib, frontend
dev = torch.device('cuda:0')
res = {}
for fused in (sys.argv[1:] or ["1", "0"]):
    os.environ["TALFE_FUSED_NORM"] = fused
    _lib._LIB = None; frontend._PLANS.clear()
    lib = _lib.load(); mod = LogMelSpec().to(dev)
    for B, secs in ((1, 60), (1, 1)):
        L = secs * 16000
        w = torch.empty(B, L, device=dev)
        _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, B, L, L, 2020, 0, 0, None))
        out = torch.empty(B, 1 + L // 160, 80, device=dev)
        r = {}
        for norm in ("none", "batch"):
            fn = lambda: mod.features(w, norm=norm, out=out)
            for _ in range(20): fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200): fn()
            t_host = (time.perf_counter() - t0) / 200 * 1e6        # enqueue cost (the device lags behind)
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
            r[norm] = {"host_enqueue_us": round(t_host, 2), "back_to_back_us": round(t_all, 2), "graph_us": round(e0.elapsed_time(e1) / 200 * 1e3, 2)}
        res[f"fused={fused} {B}x{secs}s"] = r
print(json.dumps(res, indent=1))
