#!/bin/bash
# resident kernel with c in tensor memory and a 5-stage Q ring: Lasso tests, then the skew sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py tests/test_edge_cases_gpu.py -x -q -k "lasso or resident or fista or ista" > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
tail -3 gpurun_out/r2ah_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista --repeats 10 > gpurun_out/r2ah_$tag.json 2>> gpurun_out/r2ah.err; }
run skew2 DECOMP_RESIDENT_SKEW=2
run skew0 DECOMP_RESIDENT_SKEW=0
run skew1 DECOMP_RESIDENT_SKEW=1
run skew3 DECOMP_RESIDENT_SKEW=3
run skew4 DECOMP_RESIDENT_SKEW=4
python - <<'PY'
import json
for n in ('skew0','skew1','skew2','skew3','skew4'):
    try:
        b=json.loads(open('gpurun_out/r2ah_%s.json'%n).read().strip().splitlines()[-1]); print(n, 'ms/step %.4f frac %.4f'%(b['ms_per_step'], b['roofline']['frac']), b['timing']['min_ms'], b['timing']['max_ms'], b['results_finite'])
    except Exception as e: print(n,'failed',e)
PY
