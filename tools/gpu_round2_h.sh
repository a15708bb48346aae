#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 300 python tools/bench_c1.py > gpurun_out/r2h_c1.json 2> gpurun_out/r2h_c1.err; cat gpurun_out/r2h_c1.json | tr -d '\n '; echo; tail -3 gpurun_out/r2h_c1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tf32|gemm_f64|reduce_partials|split_|normalize|transpose_real" --csv --log-file gpurun_out/r2h_nmf_tf32_launches.csv python tools/prof_nmf.py 1000000 2 tf32x3 > gpurun_out/r2h_ncu.log 2>&1
tail -n 2 gpurun_out/r2h_ncu.log
