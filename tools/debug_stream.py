"""Where do streamed and one-shot un-normalised features differ? (development aid)"""
import sys, json, torch
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, _lib
from tal_asrd_b200.streaming import stream_episode
dev = torch.device('cuda:0'); lib = _lib.load(); mod = LogMelSpec().to(dev)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 57_600_000
ep = torch.empty(1, L, device=dev)
_lib.check(lib.talfe_synth_fill(ep.data_ptr(), _lib.F32, 1, L, L, 2020, 42, 0, None))
ep = ep[0]
one = mod.features(ep[None], norm="none")
one2 = mod.features(ep[None], norm="none")
print("one-shot deterministic:", torch.equal(one, one2))
host = ep.cpu().pin_memory()
for name, src, cs in (("host120", host, 120.0), ("dev120", ep, 120.0), ("dev30", ep, 30.0), ("host30", host, 30.0), ("dev32f", ep, 32 * 100 / 100.0)):
    got = stream_episode(mod, src, cs, device=dev, normalise=False)
    torch.cuda.synchronize()
    d = (got != one)
    n = int(d.sum())
    print(name, "differing elements:", n)
    if n:
        idx = d.nonzero()
        fr = idx[:, 1]
        print("  frames (first 20):", fr[:20].tolist(), "mels:", idx[:20, 2].tolist())
        print("  distinct frames:", torch.unique(fr).numel(), "frame mod 3000 / 12000 of first:", int(fr[0]) % 3000, int(fr[0]) % 12000)
        print("  max abs diff:", float((got - one).abs().max()))
        uf = torch.unique(fr)
        print("  unique frames sample:", uf[:40].tolist())
