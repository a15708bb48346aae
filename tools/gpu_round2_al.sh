#!/bin/bash
# refresh after the TN tile-order change: masked FP64 sweep traffic, then the driver's bench command
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2al_masked_fp64_sweep.csv python tools/prof_nmf.py 1000000 3 fp64 1024 128 1 > gpurun_out/r2al_ncu.log 2>&1
python tools/sweep_traffic.py gpurun_out/r2al_masked_fp64_sweep.csv > gpurun_out/r2al_masked_fp64_traffic.json
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2al_masked_fp64_traffic.json')); print('masked fp64 sweep bytes %.4g ms %.2f'%(j['dram_bytes'], j['kernel_time_ms']))
d=json.loads(open('gpurun_out/r2al_bench.json').read().strip().splitlines()[-1])
print('fista', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_call'], 'nmf', d['secondary']['ms_per_step'], d['secondary']['roofline']['frac'], 'c5', d['extra_configs']['c5_masked_nmf_sweep']['ms_per_step'], 'c4', d['extra_configs']['c4_dl_masked_step']['ms_per_step'], d.get('leg_errors'))
PY
