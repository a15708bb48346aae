#!/bin/bash
# first GPU pass of round 2: tests, measured parity errors, the driver's bench command for both arms, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python tools/parity_errors.py > gpurun_out/r2a_parity_errors.json 2> gpurun_out/r2a_parity_errors.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 20 --warmup 5 --repeats 2 --legs fista,nmf --rows 131072 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; tail -c 600 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
