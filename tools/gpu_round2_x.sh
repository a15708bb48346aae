#!/bin/bash
# rows-per-CTA sweep of the one-launch small NMF kernel at configs[0]
mkdir -p gpurun_out
for r in 0 8 10 12 14 20 28; do
  DECOMP_SMALL_ROWS=$r python tools/bench_c1.py > gpurun_out/r2x_c1_rows$r.json 2> gpurun_out/r2x_err.log || tail -3 gpurun_out/r2x_err.log
  python - <<PY
import json
d=json.load(open('gpurun_out/r2x_c1_rows$r.json'))
n=d['nomask']; m=d['mask']
print('rows $r: nomask %.3f ms (%.0f sweeps/s, %.1f us/sweep, 1-sweep call %.3f ms)  mask %.3f ms (%.1f us/sweep)  errD %.1e' % (n['gpu_s']*1e3, n['gpu_sweeps_per_s'], n['us_per_sweep'], n['call_1_sweep_s']*1e3, m['gpu_s']*1e3, m['us_per_sweep'], n['rel_err_D']))
PY
done
