#!/bin/bash
# One GPU-box visit: GPU test suite on the shipped build, A/B against every experiment build under tools/_abl/
# (tools/build_variant.sh; ab_libs.py also names the fastest bit-identical one in gpurun_out/cand.txt), bench line,
# ncu launch list + full capture of K1, the reference arm.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
t0=$(date +%s)
mark() { echo "$1 done $(( $(date +%s) - t0 )) s" >> gpurun_out/timeline.txt; }
timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -25 > gpurun_out/pytest_default.log; mark "pytest"
timeout 200 python tools/ab_libs.py gpurun_out/ab_libs.json tal_asrd_b200/libtalfe.so $(ls tools/_abl/*.so 2>/dev/null) > gpurun_out/ab_libs.log 2>&1; mark "ab"
timeout 240 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; mark "bench"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_default.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:logmel -s 3 -c 1 -f -o gpurun_out/prof_default \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_default.log 2>&1; mark "ncu"
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; mark "reference arm"
[ -n "$EXTRA" ] && { timeout 200 bash -c "$EXTRA" > gpurun_out/extra.log 2>&1; mark "extra"; }
tail -4 gpurun_out/pytest_default.log; cat gpurun_out/bench_default.json; cat gpurun_out/timeline.txt
