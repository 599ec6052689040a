#!/bin/bash
# quick iteration on the GPU box: parity tests, A/B of the kernel variants, bench line, ncu launch list + full capture
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/pytest.log
timeout 120 python tools/ab_kernels.py 30 > gpurun_out/ab.json 2> gpurun_out/ab.err
timeout 300 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "${PROFILE:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:logmel -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
fi
tail -15 gpurun_out/pytest.log; cat gpurun_out/ab.json; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ab.err
