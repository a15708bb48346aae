#!/bin/bash
# 8-GPU pass: the driver's scaling command at N=8 for both arms (and N=4 for ours), C1 timing on one GPU
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2p_gpus.txt
timeout 300 python tools/bench_c1.py > gpurun_out/r2p_c1.json 2> gpurun_out/r2p_c1.err; cat gpurun_out/r2p_c1.json | tr -d '\n '; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p_bench8_ref.json 2> gpurun_out/r2p_bench8_ref.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p_bench8.json 2> gpurun_out/r2p_bench8.err; echo "rc=$?" >> gpurun_out/r2p_bench8.err
tail -3 gpurun_out/r2p_bench8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 --legs fista,e2e,nmf_strong,tf32,parity > gpurun_out/r2p_bench4.json 2> gpurun_out/r2p_bench4.err; echo "rc=$?" >> gpurun_out/r2p_bench4.err
python - <<'PY'
import json
for n in ('r2p_bench8','r2p_bench4'):
    try:
        b=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'fista %.4g frac %.3f'%(b['value'], b['roofline']['frac']), 'e2e ms', b['e2e']['ms_per_call'], 'pinned', b['e2e_pinned']['ms_per_call'], 'h2d', b['e2e']['h2d_copy_only']['ms'], b['e2e']['h2d_copy_only']['gbs_all_gpus'])
        s=b['secondary_strong']; print('  strong ms', s['ms_per_step'], 'allreduce', s['allreduce_ms_per_sweep'], 'tf32', s.get('tf32x3',{}).get('ms_per_step'))
        if 'secondary' in b: print('  weak ms', b['secondary']['ms_per_step'], b['secondary'].get('tf32x3',{}).get('ms_per_step'))
        print('  parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'])
        for k,v in b.get('extra_configs',{}).items():
            print('  ', k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -c 400 gpurun_out/r2p_bench8_ref.json
