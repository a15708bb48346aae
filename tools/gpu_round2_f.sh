#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_nmf_tf32_launches.csv python tools/prof_nmf.py 262144 3 tf32x3 > gpurun_out/r2f_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 4 -c 2 -o gpurun_out/r2f_tf32 python tools/prof_nmf.py 262144 3 tf32x3 > gpurun_out/r2f_ncu2.log 2>&1
tail -2 gpurun_out/r2f_ncu.log gpurun_out/r2f_ncu2.log
