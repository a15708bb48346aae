#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -x -q -k "nmf" > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
python tools/bench_c1.py > gpurun_out/r2z_c1.json 2> gpurun_out/r2z_err.log || tail -3 gpurun_out/r2z_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/r2z_c1.json'))
n=d['nomask']; m=d['mask']
print('nomask %.3f ms (%.0f sweeps/s, %.1f us/sweep, 1-sweep call %.3f ms)  mask %.3f ms (%.0f sweeps/s, %.1f us/sweep)  errD %.1e %.1e' % (n['gpu_s']*1e3, n['gpu_sweeps_per_s'], n['us_per_sweep'], n['call_1_sweep_s']*1e3, m['gpu_s']*1e3, m['gpu_sweeps_per_s'], m['us_per_sweep'], n['rel_err_D'], m['rel_err_D']))
PY
