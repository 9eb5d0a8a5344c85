#!/bin/bash
# compute-sanitizer memcheck over the kernels added / rewritten in round 2 (small test cases only; run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash scripts/sanitize_new_kernels.sh'
mkdir -p gpurun_out
K="sgemm_f32 or conv2d_f32 or dwconv_f32 or litemla_attn_f32 or bilinear_nhwc_f32 or decoder_twins or wgrad_tc or pw_small or colsum or layernorm_bf16 or win_attn or litemla_attn_bwd or bilinear_bwd or stem_wgrad or test_gemm_tc or gemm_strided"
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 \
  python -m pytest tests/test_strict_gpu.py tests/test_ops_gpu.py tests/test_zz_train_gpu.py -m gpu -x -q -p no:cacheprovider -k "$K" \
  > gpurun_out/r2_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2_sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2_sanitizer.log | tail -5
