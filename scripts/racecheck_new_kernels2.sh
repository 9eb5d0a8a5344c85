#!/bin/bash
# compute-sanitizer racecheck over the shared-memory kernels of the second half of round 2 that synchronise with plain barriers /
# cp.async groups (dw_tc double buffering, dw_wgrad_win tiles, cp.async ring reductions, dw_tiled); small cases, ONE GPU, under gpurun.
mkdir -p gpurun_out
K="dwconv_tc or dwconv_wgrad_win or test_dwconv or bn_stats or bn_act"
timeout 700 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 86 --print-limit 10 \
  python -m pytest tests/test_ops_gpu.py tests/test_zz_train_gpu.py -m gpu -x -q -p no:cacheprovider -k "$K" \
  > gpurun_out/r2b_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2b_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|rc=|hazard" gpurun_out/r2b_racecheck.log | tail -8
