#!/bin/bash
# One GPU-box visit: full GPU test suite on the default build, A/B of the experiment builds under tools/_abl/ (which
# also picks the fastest bit-identical build as the candidate), bench line, then tests + bench for the candidate,
# ncu launch list + full capture of K1 for both, and the reference arm.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
A=tools/_abl
t0=$(date +%s)
mark() { echo "$1 done $(( $(date +%s) - t0 )) s" >> gpurun_out/timeline.txt; }
timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -25 > gpurun_out/pytest_default.log; mark "pytest default"
timeout 170 python tools/ab_libs.py gpurun_out/ab_libs.json tal_asrd_b200/libtalfe.so $A/libtalfe_v3base.so $A/libtalfe_wconst.so \
    $A/libtalfe_twreg.so $A/libtalfe_wc_tw.so $A/libtalfe_skel_packed.so $A/libtalfe_namedbar.so $A/libtalfe_both.so $A/libtalfe_nb_tw.so \
    $A/libtalfe_all3.so $A/libtalfe_skel_all3.so > gpurun_out/ab_libs.log 2>&1; mark "ab"
timeout 240 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; mark "bench default"
CAND=$(cat gpurun_out/cand.txt 2>/dev/null)
if [ -n "$CAND" ]; then
  TALFE_LIB=$PWD/$CAND timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -25 > gpurun_out/pytest_cand.log; mark "pytest cand"
  TALFE_LIB=$PWD/$CAND timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cand.json 2> gpurun_out/bench_cand.err; mark "bench cand"
fi
for v in cand default; do
  if [ $v = cand ]; then [ -n "$CAND" ] || continue; export TALFE_LIB=$PWD/$CAND; else unset TALFE_LIB; fi
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$v.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$v.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:logmel -s 3 -c 1 -f -o gpurun_out/prof_$v \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$v.log 2>&1
  mark "ncu $v"
done
unset TALFE_LIB
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; mark "reference arm"
tail -4 gpurun_out/pytest_default.log; tail -4 gpurun_out/pytest_cand.log; tail -12 gpurun_out/ab_libs.json; cat gpurun_out/timeline.txt
