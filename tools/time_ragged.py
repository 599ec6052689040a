"""configs[3] (64 rows, 1 s .. 10 min, log-uniform): padded / per-row / packed calls, CUDA events, ms per call."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(2020)
B = 64
lens = torch.exp(torch.rand(B, generator=g) * (math.log(9_600_000) - math.log(16_000)) + math.log(16_000)).long()
lens[0] = 9_600_000
L = int(lens.max())
x = torch.empty(B, L, dtype=torch.float32, device=dev)
_lib.check(lib.talfe_synth_fill(x.data_ptr(), _lib.F32, B, L, L, 2020, 3, 0, None))
mod = LogMelSpec().to(dev)
lens_d = lens.to(dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
valid = int((1 + lens // 160).sum())
print("valid frames", valid, "padded frames", B * (1 + L // 160))
out = torch.empty(B, 1 + L // 160, 80, dtype=torch.float32, device=dev)
print("padded (reference) ms", round(t(lambda: mod.features(x, out=out)), 4))
xz = x.clone()
for r in range(B):
    xz[r, int(lens[r]):] = 0
print("padded, lens as padding hint ms", round(t(lambda: mod.features(xz, audio_lens=lens_d, lens_are_padding=True, out=out)), 4), " (plain call on the same zero-padded buffer:", round(t(lambda: mod.features(xz, out=out)), 4), ")")
fast = LogMelSpec(detect_padding=True).to(dev)
print("drop-in forward(), detect_padding=True ms", round(t(lambda: fast(xz)), 4), " (scan alone:", round(t(lambda: fast.padding_lens(xz)), 4), ")")
dense = torch.empty(64, 480000, dtype=torch.float32, device=dev)
_lib.check(lib.talfe_synth_fill(dense.data_ptr(), _lib.F32, 64, 480000, 480000, 2020, 9, 0, None))
print("dense 64 x 30 s: plain forward ms", round(t(lambda: mod(dense), 20), 4), " with detect_padding ms", round(t(lambda: fast(dense), 20), 4))
print("per-row            ms", round(t(lambda: mod.features(x, audio_lens=lens_d, norm="row", out=out)), 4))
print("packed row         ms", round(t(lambda: mod.features_packed(x, lens, norm="row")), 4))
print("packed row_mel_var ms", round(t(lambda: mod.features_packed(x, lens, norm="row_mel_var")), 4))
