#!/bin/bash
# ncu --set full (with source) over ONE EV-M forward at the bench shape; exports CSVs and drops the (>64 MiB) report.
# usage: bash scripts/ncu_fwd_capture.sh <tag> [model]
tag=${1:-r2}
model=${2:-efficientvit_b1}
out=gpurun_out
mkdir -p $out
rep=/tmp/${tag}_fwd
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:es3:: --launch-skip 186 --launch-count 62 \
    -o $rep -f python scripts/evm_once.py 32 $model > $out/${tag}_ncu.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > $out/${tag}_fwd_raw.csv 2>>$out/${tag}_ncu.log
for k in mbconv_tc_kernel mbconv_tc_s2_kernel litemla_aggreg_dwpw stem_fused dwproj_tc litemla_kv_tc litemla_apply_tc bilinear gemm_tc_kernel; do
  ncu -i $rep.ncu-rep --page source --csv --kernel-name-base demangled -k regex:$k -c 1 > $out/${tag}_src_$k.csv 2>>$out/${tag}_ncu.log
done
gzip -f $out/${tag}_src_*.csv $out/${tag}_fwd_raw.csv
ls -la $out
