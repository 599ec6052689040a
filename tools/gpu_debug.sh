#!/bin/bash
# bounded debugging pass: every step under its own short timeout so that a hung kernel cannot eat the GPU budget
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/dbg_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/dbg_smoke.log
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -v 2>&1 | tail -30 > gpurun_out/dbg_parity.log; echo "parity rc=${PIPESTATUS[0]}"
tail -12 gpurun_out/dbg_parity.log
