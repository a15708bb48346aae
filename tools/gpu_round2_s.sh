#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista --repeats 10 > gpurun_out/r2s_$tag.json 2>> gpurun_out/r2s.err; }
run base DECOMP_RESIDENT_SKEW=2
run skew0 DECOMP_RESIDENT_SKEW=0
run skew1 DECOMP_RESIDENT_SKEW=1
run nopf DECOMP_RESIDENT_PREFETCH=0
run base32 DECOMP_RESIDENT_SKEW=2
python - <<'PY'
import json
for n in ('base','skew0','skew1','nopf','base32'):
    try:
        b=json.load(open('gpurun_out/r2s_%s.json'%n)); print(n, 'ms/step %.4f frac %.4f'%(b['ms_per_step'], b['roofline']['frac']), b['timing']['min_ms'], b['timing']['max_ms'])
    except Exception as e: print(n,'failed',e)
PY
timeout 300 python bench.py --gpus 1 --steps 32 --warmup 5 --legs fista --repeats 10 > gpurun_out/r2s_steps32.json 2>> gpurun_out/r2s.err
python -c "
import json; b=json.load(open('gpurun_out/r2s_steps32.json')); print('steps32 ms/step %.4f frac %.4f'%(b['ms_per_step'], b['roofline']['frac']))"
