#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -3 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2q_bench_nmf.json 2> gpurun_out/r2q_bench_nmf.err; echo "rc=$?" >> gpurun_out/r2q_bench_nmf.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2q_bench_nmf.json')); t=b['tf32x3']
    print('tf32 ms', t['ms_per_step'], 'frac', t['roofline']['frac'], 'err', t['max_rel_diff_D_vs_fp64'], 'fp64', b['ms_per_step'])
except Exception as e:
    print('failed', e)
PY
