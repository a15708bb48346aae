#!/bin/bash
# K-chunked y: tests + C3 TF32 sweep launch lists with and without
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tf32x3_gpu.py -x -q > gpurun_out/r2ab_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ab_pytest.log
tail -4 gpurun_out/r2ab_pytest.log
for ch in 1 0; do
DECOMP_TF32_Y_CHUNKED=$ch ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_c3_$ch.csv \
  python tools/prof_nmf.py 524288 3 tf32x3 > gpurun_out/r2ab_ncu.log 2>&1
DECOMP_TF32_Y_CHUNKED=$ch ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_c5_$ch.csv \
  python tools/prof_nmf.py 1000000 3 tf32x3 1024 128 1 > gpurun_out/r2ab_ncu.log 2>&1
python - <<PY
import csv
for tag in ('c3', 'c5'):
    rows = list(csv.reader(open('gpurun_out/r2ab_%s_$ch.csv' % tag)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
    names = [s[0] for s in seq]
    idx = [i for i, n in enumerate(names) if 'normalize' in n]
    lo, hi = idx[-2] + 1, idx[-1] + 1
    tot = sum(v for _, v in seq[lo:hi])
    print(tag, 'chunked=$ch sweep %.3f ms:' % (tot / 1e6), ' '.join('%s=%.2f' % (n.split('<')[1][:1] if '<' in n else n[5:12], v / 1e6) for n, v in seq[lo:hi] if v > 2e5))
PY
done
