"""HostPipeline on the headline batch (64 x 30 s int16 PCM in, float32 features out): ms per step by pipeline depth and number
of timed steps (ramp-up and drain amortise over the run)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, HostPipeline
dev = torch.device("cuda:0")
B, L = 64, 480000
T = 1 + L // 160
mod = LogMelSpec().to(dev)
host_in = [torch.randint(-3000, 3000, (B, L), dtype=torch.int16).pin_memory() for _ in range(3)]
host_out = [torch.empty(B, T, 80, dtype=torch.float32).pin_memory() for _ in range(3)]
for depth in (2, 3):
    for steps in (20, 50, 100):
        pipe = HostPipeline(mod, dev, depth=depth)
        for i in range(4):
            pipe.submit(host_in[i % 3], host_out[i % 3])
        pipe.drain()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(pipe.h2d)
        for i in range(steps):
            pipe.submit(host_in[i % 3], host_out[i % 3])
        p1.record(pipe.d2h)
        pipe.drain()
        print(f"depth {depth} steps {steps:4d}: {p0.elapsed_time(p1) / steps:.4f} ms per step")
