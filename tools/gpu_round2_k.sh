#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "nmf" > gpurun_out/r2k_pytest_nmf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_nmf.log
tail -6 gpurun_out/r2k_pytest_nmf.log
timeout 300 python tools/bench_c1.py > gpurun_out/r2k_c1.json 2> gpurun_out/r2k_c1.err; cat gpurun_out/r2k_c1.json | tr -d '\n '; echo; tail -3 gpurun_out/r2k_c1.err
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
