#!/bin/bash
# Round-2 first GPU call for the training path (run under gpurun, ONE GPU; numbers printed under ncu are never bench values):
#   gpurun --timeout 600 -- 'bash scripts/ncu_train.sh'
# 1. launch list of one EV-M KD training iteration (batch 8 x 512^2 keeps the ~650 serialised launches short)
# 2. `--set full` captures of the three kernels profiles/r1_next_steps.md names first
# Read back here with scripts/agg_launches.py / scripts/ncu_summary.py and commit the summaries under profiles/.
set -x
mkdir -p gpurun_out
ARGS="scripts/train_step_bench.py --batch 8 --img 512 --embed 32 --steps 1 --warmup 1"
# (a normal, un-profiled run with the CPU baseline beside it:  python scripts/train_step_bench.py --cpu-baseline --table gpurun_out/train_step.md)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/train_launches.csv python $ARGS > gpurun_out/train_ncu.log 2>&1
for k in wgrad_pw_kernel dw_wgrad_strip_kernel col_reduce_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 2 -o gpurun_out/train_$k -f python $ARGS >> gpurun_out/train_ncu.log 2>&1
done
# the never-run code of round 1 (RepViT training graph, tiled depthwise wgrad, batched SqueezeExcite backward)
ES3_ISOLATED=1 python -m pytest tests/test_zz_train_gpu.py -q -m gpu -s -p no:cacheprovider -k "repvit or tiled or batched or generic or b2 or segmenter or layernorm_bwd or win_attn_bias_bwd or tinyvit" > gpurun_out/train_unverified.log 2>&1
tail -5 gpurun_out/train_unverified.log
