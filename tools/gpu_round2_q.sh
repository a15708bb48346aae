#!/bin/bash
# final 8-GPU pass: the driver's scaling command at N=8 for both arms, then N=2 for ours
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2q_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench8_ref.json 2> gpurun_out/r2q_bench8_ref.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench8.json 2> gpurun_out/r2q_bench8.err; echo "rc=$?" >> gpurun_out/r2q_bench8.err
tail -3 gpurun_out/r2q_bench8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err; echo "rc=$?" >> gpurun_out/r2q_bench2.err
python - <<'PY'
import json
for n in ('r2q_bench8','r2q_bench2'):
    try:
        b=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'fista %.4g frac %.3f'%(b['value'], b['roofline']['frac']), 'e2e ms', b['e2e']['ms_per_call'], 'pinned', b['e2e_pinned']['ms_per_call'], 'h2d', b['e2e']['h2d_copy_only']['ms'], b['e2e']['h2d_copy_only']['gbs_all_gpus'])
        s=b['secondary_strong']; print('  strong ms', s['ms_per_step'], 'allreduce', s['allreduce_ms_per_sweep'], 'tf32', s.get('tf32x3',{}).get('ms_per_step'))
        if 'secondary' in b: print('  weak ms', b['secondary']['ms_per_step'], b['secondary'].get('tf32x3',{}).get('ms_per_step'))
        print('  parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'])
        for k,v in b.get('extra_configs',{}).items():
            print('  ', k, 'ms', v.get('ms_per_step', v.get('ms_per_call')), 'frac', v.get('roofline',{}).get('frac'), 'tf32', v.get('tf32x3',{}).get('ms_per_step'))
        print('  errors', b.get('errors'))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -c 300 gpurun_out/r2q_bench8_ref.json
