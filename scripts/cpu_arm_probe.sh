#!/bin/bash
# Why does the CPU arm read 1.3 img/s in a fresh process and 5 img/s at the end of the native run?  (diagnosis only)
mkdir -p gpurun_out
L=gpurun_out/r2q_cpu_probe.log
: > $L
run() { echo "== $1" >> $L; shift; ( "$@" python bench.py --impl reference --steps 6 --warmup 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['cpu_baseline']['min_median_max'], d['wall_s'])" ) >> $L 2>&1; }
cat /sys/kernel/mm/transparent_hugepage/enabled >> $L 2>&1
nproc >> $L; python -c "import os,psutil; print(len(os.sched_getaffinity(0)), psutil.cpu_count(False), psutil.cpu_count(True))" >> $L
run "default" env
run "default again" env
run "MALLOC_MMAP_MAX_=0 TRIM=max" env MALLOC_MMAP_MAX_=0 MALLOC_TRIM_THRESHOLD_=68719476736 MALLOC_TOP_PAD_=1073741824
run "OMP_PROC_BIND=close OMP_PLACES=cores" env OMP_PROC_BIND=close OMP_PLACES=cores
run "KMP/GOMP spin: OMP_WAIT_POLICY=active" env OMP_WAIT_POLICY=active
run "32 threads (ES3_CPU_THREADS)" env ES3_CPU_THREADS=32
run "cuda init first (ES3_CPU_CUDA_INIT=1)" env ES3_CPU_CUDA_INIT=1
cat $L
