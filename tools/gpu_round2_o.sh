#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_tf32x3_gpu.py -m gpu -x -q -k "nmf" > gpurun_out/r2o_pytest_nmf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_nmf.log
tail -3 gpurun_out/r2o_pytest_nmf.log
timeout 300 python tools/bench_c1.py > gpurun_out/r2o_c1.json 2> gpurun_out/r2o_c1.err; cat gpurun_out/r2o_c1.json | tr -d '\n '; echo; tail -2 gpurun_out/r2o_c1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmf_mu_small -c 2 -o gpurun_out/r2o_small python tools/prof_c1.py > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log
