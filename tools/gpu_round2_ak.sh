#!/bin/bash
# m-fast tile order of the TN GEMM: tests, then DRAM traffic and time of the FP64 NMF sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_full_size_gpu.py -x -q > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ak_pytest.log
tail -3 gpurun_out/r2ak_pytest.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2ak_nmf_fp64_sweep.csv python tools/prof_nmf.py 1000000 3 fp64 > gpurun_out/r2ak_ncu.log 2>&1
python tools/sweep_traffic.py gpurun_out/r2ak_nmf_fp64_sweep.csv > gpurun_out/r2ak_nmf_fp64_traffic.json
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2ak_nmf_fp64_traffic.json')); print('fp64 sweep bytes %.4g ms %.2f'%(j['dram_bytes'], j['kernel_time_ms']), [(k['name'][-24:], round(k['ms'],2), round((k['read']+k['written'])/1e9,1)) for k in j['kernels']])
PY
timeout 600 python bench.py --legs nmf --nmf-steps 4 > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ak_bench.json').read().strip().splitlines()[-1])
print('nmf ms', d['ms_per_step'], d['roofline']['frac'], d.get('leg_errors'))
PY
