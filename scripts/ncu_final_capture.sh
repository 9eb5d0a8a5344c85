#!/bin/bash
# Final-round ncu evidence (ONE GPU, under gpurun): `--set full` raw metrics of
#   (a) one EV-M forward at the bench shape (62 launches),
#   (b) the training step's backward kernels (a 512^2 / batch-8 step keeps the capture short),
#   (c) the teacher's GEMMs and attention kernels (batch 2).
# Exports raw CSVs only (reports are too large for gpurun_out); scripts/ncu_table.py renders them into profiles/*.md.
set -x
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --kernel-name-base demangled -k regex:es3:: --launch-skip 186 --launch-count 62 -o /tmp/r2s_fwd -f \
    python scripts/evm_once.py 32 efficientvit_b1 > $out/r2s_ncu.log 2>&1
ncu -i /tmp/r2s_fwd.ncu-rep --page raw --csv 2>>$out/r2s_ncu.log | gzip > $out/r2s_fwd_raw.csv.gz
ncu --set full --clock-control none --kernel-name-base demangled \
    -k regex:"wgrad_tc_kernel|wgrad_tc_reduce|col_reduce_kernel|bn_act_bwd_apply|litemla_dqkv|litemla_dkv|dw_bwd_data_s2k3|stem_wgrad_kernel|dw_wgrad_tiled|bilinear_bwd|pw_small" \
    --launch-skip 200 --launch-count 60 -o /tmp/r2s_train -f \
    python scripts/train_step_bench.py --batch 8 --img 512 --embed 32 --steps 1 --warmup 1 >> $out/r2s_ncu.log 2>&1
ncu -i /tmp/r2s_train.ncu-rep --page raw --csv 2>>$out/r2s_ncu.log | gzip > $out/r2s_train_raw.csv.gz
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"attn_win24|attn_tc_kernel|gemm_tc_kernel" \
    --launch-skip 60 --launch-count 14 -o /tmp/r2s_teacher -f python scripts/family_table.py teacher 2 1008 >> $out/r2s_ncu.log 2>&1
ncu -i /tmp/r2s_teacher.ncu-rep --page raw --csv 2>>$out/r2s_ncu.log | gzip > $out/r2s_teacher_raw.csv.gz
ls -la $out | grep r2s
