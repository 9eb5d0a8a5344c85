#!/bin/bash
# ncu evidence of the code at the end of round 2 (ONE GPU, under gpurun): `--set full` raw metrics of
#   (a) one EV-M forward at the bench shape (launches of the 4th forward; the tap-rounding kernels run in the first only),
#   (b) the depthwise kernels at the training shapes (scripts/dw_bench.py): dw_tc_kernel, dw_tiled_kernel,
#   (c) the sliding-window depthwise weight gradient and the column reductions inside a small training step.
# Raw CSVs only; scripts/ncu_table.py renders them.
set -x
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --kernel-name-base demangled -k regex:es3:: --launch-skip 200 --launch-count 62 -o /tmp/r2ag_fwd -f \
    python scripts/evm_once.py 32 efficientvit_b1 > $out/r2ag_ncu.log 2>&1
ncu -i /tmp/r2ag_fwd.ncu-rep --page raw --csv 2>>$out/r2ag_ncu.log | gzip > $out/r2ag_fwd_raw.csv.gz
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"dw_tc_kernel|dw_tiled_kernel" --launch-skip 3 --launch-count 1 -o /tmp/r2ag_dw1 -f \
    python scripts/dw_bench.py tc >> $out/r2ag_ncu.log 2>&1
ncu -i /tmp/r2ag_dw1.ncu-rep --page raw --csv 2>>$out/r2ag_ncu.log | gzip > $out/r2ag_dwtc_raw.csv.gz
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"dw_wgrad_win_kernel|col_reduce_kernel|dw_tc_kernel" \
    --launch-skip 150 --launch-count 40 -o /tmp/r2ag_train -f \
    python scripts/train_step_bench.py --batch 8 --img 512 --embed 32 --steps 1 --warmup 1 >> $out/r2ag_ncu.log 2>&1
ncu -i /tmp/r2ag_train.ncu-rep --page raw --csv 2>>$out/r2ag_ncu.log | gzip > $out/r2ag_train_raw.csv.gz
ls -la $out | grep r2ag
