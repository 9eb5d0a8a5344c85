#!/bin/bash
# compute-sanitizer memcheck over the kernels of the second half of round 2 (small test cases only; ONE GPU, under gpurun):
# tensor-core depthwise conv, tap rounding, sliding-window depthwise weight gradients, fused MBConv / dw+project / aggregation / stem
# with shared fragments and two-tap MMAs, stride-2 plane layout, cp.async ring reductions, strict attention / RoPE / LayerNorm.
mkdir -p gpurun_out
K="dwconv_tc or dwconv_wgrad_win or test_dwconv or mbconv or dwproj or aggreg or stem_fused or bn_stats or bn_act or attention_f32 or rope_and_scale or vit_trunk"
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 \
  python -m pytest tests/test_strict_gpu.py tests/test_ops_gpu.py tests/test_zz_train_gpu.py -m gpu -x -q -p no:cacheprovider -k "$K" \
  > gpurun_out/r2b_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2b_sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|out of bounds" gpurun_out/r2b_sanitizer.log | tail -8
