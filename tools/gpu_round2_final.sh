#!/bin/bash
# final measurement pass of round 2 (one GPU): tests, both bench arms, parity errors, launch lists, DRAM traffic per
# sweep, ncu --set full of the dominant kernels
mkdir -p gpurun_out
O=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
timeout 600 python tools/parity_errors.py > ${O}_parity_errors.json 2> ${O}_parity_errors.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?" >> ${O}_bench.err
tail -2 ${O}_bench.err
timeout 300 python tools/bench_c1.py > ${O}_c1.json 2> ${O}_c1.err
# launch list of the driver's command (FISTA legs)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lasso_resident|gemm_f64|tf32x3|proxq|masked_b2b" --csv --log-file ${O}_bench_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,tf32 > ${O}_ncu_launches.log 2>&1
# per-sweep launch lists with DRAM traffic: FP64 / TF32 NMF at configs[2], masked FP64 / TF32 at configs[4]
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file ${O}_nmf_fp64_sweep.csv python tools/prof_nmf.py 1000000 3 fp64 > ${O}_ncu_a.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file ${O}_nmf_tf32_sweep.csv python tools/prof_nmf.py 1000000 3 tf32x3 > ${O}_ncu_b.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file ${O}_masked_fp64_sweep.csv python tools/prof_nmf.py 1000000 3 fp64 1024 128 1 > ${O}_ncu_c.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file ${O}_masked_tf32_sweep.csv python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > ${O}_ncu_d.log 2>&1
for t in nmf_fp64 nmf_tf32 masked_fp64 masked_tf32; do python tools/sweep_traffic.py ${O}_${t}_sweep.csv > ${O}_${t}_traffic.json 2>> ${O}_traffic.err; done
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lasso_resident -s 2 -c 1 -f -o ${O}_resident python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista --repeats 2 > ${O}_ncu_res.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 7 -c 2 -f -o ${O}_tf32_nmf python tools/prof_nmf.py 262144 3 tf32x3 > ${O}_ncu_tf32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_pair_kernel -s 4 -c 4 -f -o ${O}_tf32_masked python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > ${O}_ncu_tf32m.log 2>&1
ls -la gpurun_out/r2f*.ncu-rep
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('fista', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e'].get('ms_per_call'), 'errors', d.get('errors'))
PY
