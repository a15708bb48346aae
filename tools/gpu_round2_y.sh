#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_round2_x.sh
python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > gpurun_out/r2y_plain.log 2>&1 || { tail -5 gpurun_out/r2y_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_pair_kernel -s 4 -c 1 \
  -o gpurun_out/r2y_fmask -f python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > gpurun_out/r2y_ncu.log 2>&1
tail -2 gpurun_out/r2y_ncu.log
