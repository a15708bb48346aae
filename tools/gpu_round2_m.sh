#!/bin/bash
# measurement pass: full tests, the driver's bench command for both arms, parity errors, launch lists, ncu --set full
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
timeout 600 python tools/parity_errors.py > gpurun_out/r2m_parity_errors.json 2> gpurun_out/r2m_parity_errors.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?" >> gpurun_out/r2m_bench.err
tail -2 gpurun_out/r2m_bench.err
timeout 300 python tools/bench_c1.py > gpurun_out/r2m_c1.json 2> gpurun_out/r2m_c1.err; cat gpurun_out/r2m_c1.json | tr -d '\n '; echo
# launch list of the same command (fista legs), per-launch durations
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lasso_resident|gemm_f64|tf32x3|proxq|masked_b2b" --csv --log-file gpurun_out/r2m_bench_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,tf32 > gpurun_out/r2m_ncu_launches.log 2>&1
# full captures: resident kernel in the driver's configuration (20-iteration launch), fused masked kernel, tf32 x update
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lasso_resident -s 2 -c 1 -o gpurun_out/r2m_resident python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista --repeats 2 > gpurun_out/r2m_ncu_res.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:masked_b2b -s 4 -c 1 -o gpurun_out/r2m_b2b python bench.py --gpus 1 --legs configs --c5-rows 262144 > gpurun_out/r2m_ncu_b2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 5 -c 2 -o gpurun_out/r2m_tf32 python tools/prof_nmf.py 262144 3 tf32x3 > gpurun_out/r2m_ncu_tf32.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
