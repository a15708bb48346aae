#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32x3_gpu.py -m gpu -x -q -s > gpurun_out/r2d_pytest_tf32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_tf32.log
grep -i "tf32x3\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2d_pytest_tf32.log | tail -25
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2d_bench_nmf.json 2> gpurun_out/r2d_bench_nmf.err; echo "rc=$?" >> gpurun_out/r2d_bench_nmf.err
tail -3 gpurun_out/r2d_bench_nmf.err
for kb in 512 2048; do
  DECOMP_STAGE_PIECE_KB=$kb timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e --repeats 2 > gpurun_out/r2d_e2e_p$kb.json 2>> gpurun_out/r2d_e2e.err
done
DECOMP_STAGE_PIECE_KB=1024 DECOMP_STAGE_THREADS=10 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e --repeats 2 > gpurun_out/r2d_e2e_p1024t10.json 2>> gpurun_out/r2d_e2e.err
python - <<'PY'
import json
for n in ['r2d_e2e_p512','r2d_e2e_p2048','r2d_e2e_p1024t10']:
    try:
        b=json.load(open('gpurun_out/%s.json'%n))
        print(n, 'e2e ms', b.get('e2e',{}).get('ms_per_call'), 'pinned', b.get('e2e_pinned',{}).get('ms_per_call'))
    except Exception as e:
        print(n, 'failed', e)
try:
    b=json.load(open('gpurun_out/r2d_bench_nmf.json'))
    print('nmf fp64 ms', b['ms_per_step'], 'tf32', json.dumps(b.get('tf32x3'))[:600])
except Exception as e:
    print('nmf failed', e)
PY
