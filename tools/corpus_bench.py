"""BASELINE configs[4]: a ~600-hour synthetic 16 kHz corpus sharded by episode over the ranks of one box, with ONE
all-reduce of the statistics block for dataset-level CMVN.

    python tools/corpus_bench.py                                  (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/corpus_bench.py

Pass 1 transforms every episode of this rank (un-normalised log-mel into a reused output buffer) and accumulates
{count, sum, sumsq, per-mel sums, per-mel sumsq} on the device; the blocks are all-reduced once (NCCL over NVLink);
pass 2 transforms again and applies the GLOBAL per-mel mean / variance in place.  The episodes are 1 h each
(57.6 M samples -> 360 001 frames); a pool of 8 distinct episodes per rank (1.8 GB, far larger than L2) stands in
for the rank's 600 / N episodes, which are visited in turn — the arithmetic and the traffic per episode are those of
the full corpus, only the sample values repeat.  Prints one JSON line (rank 0); times are the max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tal_asrd_b200 import LogMelSpec, _lib  # noqa: E402
from tal_asrd_b200.corpus import CorpusStats, shard_episodes  # noqa: E402

EPISODES, L, POOL = 600, 57_600_000, 8
T = 1 + L // 160
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
mod = LogMelSpec().to(dev)
mine = shard_episodes([L] * EPISODES, world, rank)
pool = []
for i in range(POOL):
    w = torch.empty(1, L, dtype=torch.float32, device=dev)
    _lib.check(lib.talfe_synth_fill(w.data_ptr(), _lib.F32, 1, L, L, 2020, rank * POOL + i, 0, None))
    pool.append(w)
out = [torch.empty(1, T, 80, dtype=torch.float32, device=dev) for _ in range(2)]
blocks = mod.stats_block(dev, rows=len(mine))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def pass1():
    for n, _ in enumerate(mine):
        mod.features(pool[n % POOL], norm="row_mel_var", stats=blocks[n:n + 1], defer_normalise=True, out=out[n % 2])
    total = CorpusStats(80, dev)
    total.add(blocks)
    return total


def pass2(total):
    for n, _ in enumerate(mine):
        y = mod.features(pool[n % POOL], norm="none", out=out[n % 2])
        mod.apply_stats(y, total.block, norm="row_mel_var")


for _ in range(2):                      # warm-up of both passes on a few episodes
    mod.features(pool[0], norm="row_mel_var", stats=blocks[:1], defer_normalise=True, out=out[0])
REPS = 3                                # the whole two-pass job, repeated; the fastest repetition (max over ranks) is reported:
best = None                             # at 40 ms per pass one descheduled host thread on one rank is visible
for rep in range(REPS):
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    total = pass1()
    ev[1].record()
    total.all_reduce()
    ev[2].record()
    pass2(total)
    ev[3].record()
    barrier()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if best is None or float(t.sum()) < float(best.sum()):
        best = t
t = best
n_local = torch.tensor([len(mine)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(n_local, op=dist.ReduceOp.SUM)
if rank == 0:
    frames = float(n_local.item()) * T
    t1, tr, t2 = [float(x) for x in t.tolist()]
    hours = float(n_local.item()) * L / 16000 / 3600
    print(json.dumps({
        "workload": f"configs[4]: {int(n_local.item())} x 1 h synthetic episodes ({hours:.0f} h), sharded by episode over {world} GPU(s), global per-mel CMVN",
        "n_gpus": world, "episodes_per_rank_max": -(-EPISODES // world), "frames": frames,
        "pass1_stats_ms": t1, "allreduce_ms": tr, "pass2_normalise_ms": t2,
        "pass1_frames_per_s": frames / (t1 * 1e-3), "pass2_frames_per_s": frames / (t2 * 1e-3),
        "both_passes_frames_per_s": frames / ((t1 + tr + t2) * 1e-3),
        "both_passes_x_realtime": hours * 3600 / ((t1 + tr + t2) * 1e-3),
        "global_mean": total.mean, "global_count": total.count,
        "repetitions": REPS, "pool": f"{POOL} distinct resident episodes per rank visited in turn (inputs 1.8 GB >> L2)"}))
if world > 1:
    dist.destroy_process_group()
