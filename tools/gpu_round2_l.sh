#!/bin/bash
# 2-GPU pass: comm wrappers, sharded solves on the native communicator, short 2-rank bench; C1 on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2l_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_dist.log
tail -5 gpurun_out/r2l_pytest_dist.log
timeout 300 python tools/bench_c1.py > gpurun_out/r2l_c1.json 2> gpurun_out/r2l_c1.err; cat gpurun_out/r2l_c1.json | tr -d '\n '; echo; tail -3 gpurun_out/r2l_c1.err
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "nmf or dictionary" > gpurun_out/r2l_pytest_nmf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_nmf.log
tail -3 gpurun_out/r2l_pytest_nmf.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --legs fista,parity,nmf_strong,configs --c5-rows 262144 > gpurun_out/r2l_bench2.json 2> gpurun_out/r2l_bench2.err; echo "rc=$?" >> gpurun_out/r2l_bench2.err
tail -3 gpurun_out/r2l_bench2.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r2l_bench2.json').read().strip().splitlines()[-1])
    s=b['secondary_strong']; print('strong', s['ms_per_step'], s['allreduce_ms_per_sweep'])
    print('parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'])
    for k,v in b.get('extra_configs',{}).items():
        print(k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'))
except Exception as e:
    print('failed', e)
PY
