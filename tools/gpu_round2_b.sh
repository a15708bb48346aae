#!/bin/bash
# GPU pass b: 16-warp resident kernel + threaded uploader
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e > gpurun_out/r2b_bench20.json 2> gpurun_out/r2b_bench20.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --legs fista > gpurun_out/r2b_bench100.json 2> gpurun_out/r2b_bench100.err
for th in 4 8; do
  DECOMP_STAGE_THREADS=$th timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e --repeats 2 > gpurun_out/r2b_e2e_t$th.json 2>> gpurun_out/r2b_bench20.err
done
DECOMP_STAGE_PIECE_MB=32 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista,e2e --repeats 2 > gpurun_out/r2b_e2e_p32.json 2>> gpurun_out/r2b_bench20.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lasso_resident -s 1 -c 1 -o gpurun_out/r2b_resident python tools/prof_fista.py 45 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_pytest.log
python - <<'PY'
import json
for n in ['r2b_bench20','r2b_bench100','r2b_e2e_t4','r2b_e2e_t8','r2b_e2e_p32']:
    try:
        b=json.load(open('gpurun_out/%s.json'%n))
        print(n, 'ms/step %.4f frac %.4f'%(b['ms_per_step'], b['roofline']['frac']), 'e2e ms', b.get('e2e',{}).get('ms_per_call'), 'pinned', b.get('e2e_pinned',{}).get('ms_per_call'))
    except Exception as e:
        print(n, 'failed', e)
PY
