#!/usr/bin/env python
"""Summarise an ncu report (raw + source pages) for the log-mel kernel: key metrics, opcode mix per frame,
stall reasons, shared-memory wavefronts per access type, hottest instructions.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [frames_per_launch]"""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 192064.0
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'sm__inst_executed.sum']
d = data[0]
print(f"kernel: {d[hdr.index('Kernel Name')][:80]}")
for k in keep:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:70s} {d[i]:>18s} {units[i]}")
src = subprocess.run(f"ncu -i {rep} --page source --csv --kernel-name regex:logmel", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hidx[0]; i1 = hidx[1] - 1 if len(hidx) > 1 else len(rows)
h = rows[i0]; col = {n: i for i, n in enumerate(h)}; ins = rows[i0 + 1:i1]
def f(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
tot_s = sum(f(r, '# Samples') for r in ins); tot_i = sum(f(r, 'Instructions Executed') for r in ins)
print(f"\nwarp-instructions/frame {tot_i / frames:.1f}   smem wavefronts/frame {sum(f(r, 'L1 Wavefronts Shared') for r in ins) / frames:.1f} "
      f"(ideal {sum(f(r, 'L1 Wavefronts Shared Ideal') for r in ins) / frames:.1f})")
ops = collections.Counter(); osamp = collections.Counter()
for r in ins:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[col['Source']]); op = m.group(2) if m else '?'
    op = op if op.startswith(('LDS', 'STS', 'LDG', 'STG')) else op.split('.')[0]
    ops[op] += f(r, 'Instructions Executed'); osamp[op] += f(r, '# Samples')
print("opcode mix (warp-inst/frame, % stall samples):")
for op, n in ops.most_common(24):
    print(f"  {op:14s} {n / frames:8.2f} {100 * osamp[op] / tot_s:6.1f}%")
st = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
print("stall reasons %:", {s[6:]: round(100 * sum(f(r, s) for r in ins) / tot_s, 1) for s in st if sum(f(r, s) for r in ins) / tot_s > 0.005})
agg = collections.defaultdict(lambda: [0, 0])
for r in ins:
    if f(r, 'L1 Wavefronts Shared') > 0:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[col['Source']]); agg[m.group(2)][0] += f(r, 'L1 Wavefronts Shared'); agg[m.group(2)][1] += f(r, 'L1 Wavefronts Shared Ideal')
print("shared wavefronts/frame by op (actual / ideal):", {k: (round(v[0] / frames, 1), round(v[1] / frames, 1)) for k, v in agg.items()})
print("hottest instructions:")
for i in sorted(sorted(range(len(ins)), key=lambda i: -f(ins[i], '# Samples'))[:16]):
    r = ins[i]; s = {x[6:]: int(f(r, x)) for x in st if f(r, x) > 0}; s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {100 * f(r, '# Samples') / tot_s:5.2f}% {r[col['Source']].strip()[:64]:64s} {s}")
