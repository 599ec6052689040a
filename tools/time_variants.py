"""Times the hot path for the input formats / layouts beyond the reference fp32 [B,T,M] case (GPU box)."""
import sys, json, torch
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, _lib
from tal_asrd_b200.streaming import stream_episode
dev = torch.device('cuda:0'); lib = _lib.load(); mod = LogMelSpec().to(dev)
B, L = 64, 480000
def fill(dtype, code, rows=B, n=L, ep=0):
    w = torch.empty(rows, n, dtype=dtype, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), code, rows, n, n, 2020, ep, 0, None)); return w
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
frames = B * (1 + L // 160)
for name, dt, code in (("f32", torch.float32, _lib.F32), ("f16", torch.float16, _lib.F16), ("i16", torch.int16, _lib.I16)):
    ws = [fill(dt, code, ep=i * B) for i in range(3)]
    out = torch.empty(B, 1 + L // 160, 80, device=dev)
    k = [0]
    def step(norm):
        k[0] += 1; return mod.features(ws[k[0] % 3], norm=norm, out=out)
    res[name] = {"ms_forward": timeit(lambda: step("batch")), "ms_kernel_only": timeit(lambda: step("none"))}
    res[name]["gframes_per_s_forward"] = frames / res[name]["ms_forward"] / 1e6
ws = [fill(torch.float32, _lib.F32, ep=i * B) for i in range(3)]
k = [0]
def mt():
    k[0] += 1; return mod.features(ws[k[0] % 3], norm="batch", layout="mt")
res["f32_layout_mt"] = {"ms_forward": timeit(mt)}
lens = torch.randint(16000, L, (B,), device=dev)
res["f32_per_row_lens_rowmean"] = {"ms_forward": timeit(lambda: mod.features(ws[0], audio_lens=lens, norm="row"))}
res["f32_row_mel_var"] = {"ms_forward": timeit(lambda: mod.features(ws[1], norm="row_mel_var"))}
ep = fill(torch.float32, _lib.F32, rows=1, n=57_600_000, ep=999)[0]
res["hour_one_shot_f32"] = {"ms": timeit(lambda: mod(ep[None]), 5)}
res["hour_streamed_30s_chunks_device_resident"] = {"ms": timeit(lambda: stream_episode(mod, ep, 30.0), 3)}
eph = ep.cpu().pin_memory()
res["hour_streamed_30s_chunks_from_pinned_host"] = {"ms": timeit(lambda: stream_episode(mod, eph, 30.0, device=dev), 3)}
res["hour_frames"] = 360001
print(json.dumps(res, indent=1))
