#!/bin/bash
# masked TF32-split NMF: kernel + solver tests, then the configs[4] legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32x3_gpu.py -x -q -s -k "mask" > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -25 gpurun_out/r2t_pytest.log
timeout 900 python bench.py --legs configs,tf32 --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
e=d.get('extra_configs',{})
c=e.get('c5_masked_nmf_sweep',{})
print('c5 nmf fp64 ms', c.get('ms_per_step'), 'frac', c.get('roofline',{}).get('frac'))
t=c.get('tf32x3',{})
print('c5 nmf tf32 ms', t.get('ms_per_step'), 'errD', t.get('max_rel_diff_D_vs_fp64'), 'roofline', json.dumps(t.get('roofline'))[:600])
print('errors', d.get('errors'))
PY
