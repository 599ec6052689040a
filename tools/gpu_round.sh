#!/bin/bash
# one GPU-box visit: tests, smoke, microbench, bench.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
(nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o /tmp/mb && timeout 120 /tmp/mb) > gpurun_out/microbench.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
