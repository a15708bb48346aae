#!/bin/bash
# ncu --set full of the four pair kernels of one masked TF32 sweep (FMASK row-major, STORE, XUPD, FMASK transposed)
mkdir -p gpurun_out
python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > gpurun_out/r2w_plain.log 2>&1 || { tail -5 gpurun_out/r2w_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_pair_kernel -s 4 -c 4 \
  -o gpurun_out/r2w_masked_tf32 -f python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > gpurun_out/r2w_ncu.log 2>&1
tail -3 gpurun_out/r2w_ncu.log
ls -la gpurun_out/r2w_masked_tf32.ncu-rep
