import torch, time, os
print(os.sched_getaffinity(0).__len__(), "cpus in affinity")
dev = torch.device("cuda:0")
n = 122880000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(61460480, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(61460480, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
dt = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D alone  {n/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms)")
dt = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H alone  {61460480/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms)")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print(f"both       {dt*1e3:.3f} ms  -> H2D {n/dt/1e9:.1f} GB/s")
