#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_tf32x3_gpu.py -m gpu -x -q -k "b2b or tf32" > gpurun_out/r2i_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_new.log
tail -6 gpurun_out/r2i_pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --legs configs > gpurun_out/r2i_bench_cfg.json 2> gpurun_out/r2i_bench_cfg.err; echo "rc=$?" >> gpurun_out/r2i_bench_cfg.err
tail -3 gpurun_out/r2i_bench_cfg.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2i_bench_cfg.json'))
    for k,v in b.get('extra_configs',{}).items():
        print(k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'), 'value', v.get('value'))
except Exception as e:
    print('failed', e)
PY
