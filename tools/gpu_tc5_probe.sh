#!/bin/bash
# The tcgen05 tensor-pipe experiment (tools/tc5_probe.cu): timing run (bounded waits + a process timeout), then the same
# binary under ncu for pipe utilisation.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tc5_probe.cu -o /tmp/tc5_probe 2> gpurun_out/tc5_probe_build.log || { tail -5 gpurun_out/tc5_probe_build.log; exit 1; }
timeout 60 /tmp/tc5_probe > gpurun_out/tc5_probe.txt 2>&1
echo "exit $?" >> gpurun_out/tc5_probe.txt
cat gpurun_out/tc5_probe.txt
grep -q "TIMED OUT\|error" gpurun_out/tc5_probe.txt && exit 0
timeout 200 ncu --clock-control none --csv --log-file gpurun_out/tc5_probe_ncu.csv \
    --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
    /tmp/tc5_probe > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/tc5_probe_ncu.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
seen = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    seen.setdefault((r[0], r[ki][:60]), {})[r[mi]] = r[vi]
done = set()
for (i, k), m in seen.items():
    if k in done: continue
    done.add(k)
    print(k, {a.split(".")[0].replace("sm__", "").replace("smsp__", ""): b for a, b in m.items()})
PY
