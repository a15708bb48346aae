#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_full_size_gpu.py -m gpu -x -q -k "lasso or resident or fista or staged or pipelin or chunk" > gpurun_out/r2c_pytest_lasso.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_lasso.log
tail -3 gpurun_out/r2c_pytest_lasso.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e > gpurun_out/r2c_bench20.json 2> gpurun_out/r2c_bench20.err
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 --legs fista > gpurun_out/r2c_bench100.json 2> gpurun_out/r2c_bench100.err
for th in 3 8; do
  DECOMP_STAGE_THREADS=$th timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e --repeats 2 > gpurun_out/r2c_e2e_t$th.json 2>> gpurun_out/r2c_bench20.err
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
python - <<'PY'
import json
for n in ['r2c_bench20','r2c_bench100','r2c_e2e_t3','r2c_e2e_t8']:
    try:
        b=json.load(open('gpurun_out/%s.json'%n))
        print(n, 'ms/step %.4f frac %.4f'%(b['ms_per_step'], b['roofline']['frac']), 'e2e ms', b.get('e2e',{}).get('ms_per_call'), 'pinned', b.get('e2e_pinned',{}).get('ms_per_call'))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -5 gpurun_out/r2c_bench20.err
