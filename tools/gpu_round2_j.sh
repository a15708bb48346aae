#!/bin/bash
# 2-GPU pass: NCCL tests of the sharded solves and a short 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2j_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_dist.log
tail -15 gpurun_out/r2j_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --legs fista,e2e,parity,nmf_strong,tf32,configs --c5-rows 262144 > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err; echo "rc=$?" >> gpurun_out/r2j_bench2.err
tail -5 gpurun_out/r2j_bench2.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r2j_bench2.json').read().strip().splitlines()[-1])
    print('fista', b['value'], b['roofline']['frac'], 'e2e', b['e2e']['ms_per_call'], 'pinned', b['e2e_pinned']['ms_per_call'])
    s=b['secondary_strong']; print('strong', s['ms_per_step'], s['allreduce_ms_per_sweep'], s.get('tf32x3',{}).get('ms_per_step'))
    print('parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'])
    for k,v in b.get('extra_configs',{}).items():
        print(k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'))
except Exception as e:
    print('failed', e)
PY
