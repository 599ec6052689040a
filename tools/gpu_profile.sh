#!/bin/bash
# ncu captures of the bench command: launch list (durations) + one full capture of the dominant kernel.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 3 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
