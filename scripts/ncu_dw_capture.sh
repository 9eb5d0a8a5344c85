#!/bin/bash
# ncu --set full (+ source) over the depthwise kernels at the training shapes (scripts/dw_bench.py); exports CSVs only.
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dw_tiled_kernel --launch-skip 3 --launch-count 1 \
    -o /tmp/r2w_dw -f python scripts/dw_bench.py > $out/r2w_ncu.log 2>&1
ncu -i /tmp/r2w_dw.ncu-rep --page raw --csv 2>>$out/r2w_ncu.log | gzip > $out/r2w_dw_raw.csv.gz
ncu -i /tmp/r2w_dw.ncu-rep --page source --csv 2>>$out/r2w_ncu.log | gzip > $out/r2w_dw_src.csv.gz
ls -la $out | grep r2w
