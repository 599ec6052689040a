#!/bin/bash
# Round-2 GPU-box visit: parity tests, A/B of library builds (tools/_abl/*.so and env variants), latency sweep, bench line.
# PROFILE=1 adds the ncu launch list and a full capture of the dominant kernel.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
t0=$(date +%s)
mark() { echo "$1 done $(( $(date +%s) - t0 )) s" >> gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout ${PYTEST_TIMEOUT:-400} python -m pytest tests -m gpu -q -x --timeout 180 ${PYTEST_ARGS} 2>&1 | tail -30 > gpurun_out/pytest.log; mark "pytest"
timeout 300 python tools/ab_libs.py gpurun_out/ab_libs.json ${AB_LIBS:-tal_asrd_b200/libtalfe.so tal_asrd_b200/libtalfe.so@TALFE_FUSED_NORM=0 $(ls tools/_abl/*.so 2>/dev/null)} > gpurun_out/ab_libs.log 2>&1; mark "ab"
timeout 200 python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err; mark "latency"
timeout 300 python tools/time_variants.py > gpurun_out/variants.json 2> gpurun_out/variants.err; mark "variants"
timeout 300 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; mark "bench"
if [ "${PROFILE:-0}" = "1" ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1; mark "ncu"
fi
[ -n "$EXTRA" ] && { timeout ${EXTRA_TIMEOUT:-300} bash -c "$EXTRA" > gpurun_out/extra.log 2>&1; mark "extra"; }
tail -6 gpurun_out/pytest.log; cat gpurun_out/ab_libs.json | head -60; cat gpurun_out/latency.json; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/timeline.txt
