#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2r_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest_dist.log
tail -4 gpurun_out/r2r_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --legs parity,configs --c5-rows 131072 > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err; echo "rc=$?" >> gpurun_out/r2r_bench2.err
tail -2 gpurun_out/r2r_bench2.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r2r_bench2.json').read().strip().splitlines()[-1])
    print('parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'], b.get('leg_errors'))
    for k,v in b.get('extra_configs',{}).items():
        print(k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'))
except Exception as e:
    print('failed', e)
PY
