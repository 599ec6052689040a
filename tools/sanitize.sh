#!/bin/bash
# compute-sanitizer passes over a small but multi-tile, multi-row problem (all kernel paths: bulk + edge tiles,
# per-row lengths, narrow dtypes, streaming).  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from tal_asrd_b200 import LogMelSpec, synth
from tal_asrd_b200.streaming import stream_episode
dev = torch.device('cuda:0')
mod = LogMelSpec().to(dev)
x = torch.from_numpy(synth.batch(1, 6, 16000 * 12)).to(dev)      # 6 rows x 1201 frames = 38 tiles/row -> every CTA loops
y = mod(x)
lens = torch.tensor([192000, 100000, 7777, 201, 150000, 64000])
z = mod.features(x, audio_lens=lens, norm='row_mel_var')
h = mod(x.half()); q = mod((x * 32767).round().to(torch.int16))
s = stream_episode(mod, x[0].cpu().pin_memory(), chunk_seconds=1.7, device=dev)
torch.cuda.synchronize()
print('ok', float(y.abs().max()), float(z.abs().max()), float((s - mod(x[:1])).abs().max()))
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
