#!/bin/bash
# FMASK experiments: what bounds the epilogue -- stores, mask loads, or neither
mkdir -p gpurun_out
for dbg in 0 1 2 3 4; do
DECOMP_TF32_DBG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aa_$dbg.csv \
  python tools/prof_nmf.py 1000000 2 tf32x3 1024 128 1 > gpurun_out/r2aa_ncu.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/r2aa_$dbg.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
fm = [v/1e6 for n, v in seq if 'pair_kernel<3>' in n]
print('dbg $dbg: FMASK launches ms', ['%.3f' % v for v in fm])
PY
done
