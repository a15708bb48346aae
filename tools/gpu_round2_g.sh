#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --legs nmf,tf32 > gpurun_out/r2g_bench_nmf.json 2> gpurun_out/r2g_bench_nmf.err; echo "rc=$?" >> gpurun_out/r2g_bench_nmf.err
tail -3 gpurun_out/r2g_bench_nmf.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2g_bench_nmf.json'))
    print('nmf fp64 ms', b['ms_per_step'], 'tf32', json.dumps(b.get('tf32x3'))[:700])
except Exception as e:
    print('nmf failed', e)
PY
timeout 300 python tools/bench_c1.py > gpurun_out/r2g_c1.json 2> gpurun_out/r2g_c1.err; cat gpurun_out/r2g_c1.json; tail -3 gpurun_out/r2g_c1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dcp" --csv --log-file gpurun_out/r2g_nmf_tf32_launches.csv python tools/prof_nmf.py 1000000 2 tf32x3 > gpurun_out/r2g_ncu.log 2>&1
tail -n 2 gpurun_out/r2g_ncu.log
