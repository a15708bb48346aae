#!/bin/bash
# launch list of the masked TF32-split NMF sweep at the configs[4] shape
mkdir -p gpurun_out
python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > gpurun_out/r2u_plain.log 2>&1 || { tail -5 gpurun_out/r2u_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_masked_tf32_launches.csv \
  python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > gpurun_out/r2u_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2u_masked_tf32_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
# last sweep only: find the last normalize kernel boundaries
names = [s[0] for s in seq]
idx = [i for i, n in enumerate(names) if 'normalize' in n]
lo, hi = idx[-2] + 1, idx[-1] + 1
tot = sum(v for _, v in seq[lo:hi])
for n, v in seq[lo:hi]:
    print('%9.3f ms  %5.1f%%  %s' % (v / 1e6, 100 * v / tot, n[:110]))
print('sweep total %.3f ms' % (tot / 1e6))
PY
