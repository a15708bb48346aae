#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32x3_gpu.py -m gpu -x -q -s > gpurun_out/r2e_pytest_tf32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_tf32.log
grep -i "tf32x3\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2e_pytest_tf32.log | tail -25
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2e_bench_nmf.json 2> gpurun_out/r2e_bench_nmf.err; echo "rc=$?" >> gpurun_out/r2e_bench_nmf.err
tail -3 gpurun_out/r2e_bench_nmf.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2e_bench_nmf.json'))
    print('nmf fp64 ms', b['ms_per_step'], 'tf32', json.dumps(b.get('tf32x3'))[:900])
except Exception as e:
    print('nmf failed', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2e_nmf_tf32_launches.csv python bench.py --legs nmf,tf32 --rows 262144 > gpurun_out/r2e_ncu.log 2>&1
