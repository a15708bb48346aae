#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tf32x3_gpu.py -m gpu -x -q -s > gpurun_out/r2n_pytest_tf32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_tf32.log
grep -i "passed\|failed\|rc=\|Error\|rel err" gpurun_out/r2n_pytest_tf32.log | tail -12
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2n_bench_nmf.json 2> gpurun_out/r2n_bench_nmf.err; echo "rc=$?" >> gpurun_out/r2n_bench_nmf.err
tail -2 gpurun_out/r2n_bench_nmf.err
DECOMP_TF32_PAIR=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2n_bench_nmf_single.json 2>> gpurun_out/r2n_bench_nmf.err
python - <<'PY'
import json
for n in ('r2n_bench_nmf','r2n_bench_nmf_single'):
    try:
        b=json.load(open('gpurun_out/%s.json'%n)); t=b['tf32x3']
        print(n, 'tf32 ms', t['ms_per_step'], 'err', t['max_rel_diff_D_vs_fp64'], 'fp64', b['ms_per_step'])
    except Exception as e:
        print(n, 'failed', e)
PY
