#!/bin/bash
# TF32-split path after the epilogue rework (8 epilogue warps, mask prefetch, n-fast tile order) + masked Lasso
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tf32x3_gpu.py -x -q -s > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
grep -v "^$" gpurun_out/r2v_pytest.log | tail -30
timeout 900 python bench.py --legs configs,tf32,nmf --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
e=d.get('extra_configs',{})
c=e.get('c5_masked_nmf_sweep',{})
print('c5 nmf fp64 ms', c.get('ms_per_step'), 'frac', c.get('roofline',{}).get('frac'))
t=c.get('tf32x3',{})
print('c5 nmf tf32 ms', t.get('ms_per_step'), 'errD', t.get('max_rel_diff_D_vs_fp64'), 'hbm frac', t.get('roofline',{}).get('frac'))
c=e.get('c5_masked_fista_iter',{})
print('c5 fista fp64 ms', c.get('ms_per_step'), 'frac', c.get('roofline',{}).get('frac'))
t=c.get('tf32x3',{})
print('c5 fista tf32 ms', t.get('ms_per_step'), 'err', t.get('max_rel_diff_x_vs_fp64'), 'hbm frac', t.get('roofline',{}).get('frac'))
s=d.get('secondary',{})
print('c3 nmf fp64 ms', s.get('ms_per_step'), 'tf32 ms', s.get('tf32x3',{}).get('ms_per_step'), 'err', s.get('tf32x3',{}).get('max_rel_diff_D_vs_fp64'))
print('errors', d.get('errors'))
PY
python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > gpurun_out/r2v_plain.log 2>&1 || { tail -5 gpurun_out/r2v_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_masked_tf32_launches.csv \
  python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > gpurun_out/r2v_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2v_masked_tf32_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
names = [s[0] for s in seq]
idx = [i for i, n in enumerate(names) if 'normalize' in n]
lo, hi = idx[-2] + 1, idx[-1] + 1
tot = sum(v for _, v in seq[lo:hi])
for n, v in seq[lo:hi]:
    print('%9.3f ms  %5.1f%%  %s' % (v / 1e6, 100 * v / tot, n[:70]))
print('sweep total %.3f ms' % (tot / 1e6))
PY
