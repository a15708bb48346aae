#!/bin/bash
# resident kernel: DRAM traffic with and without the producer's L2 prefetch of the next row block
mkdir -p gpurun_out
for pf in 1 0; do
DECOMP_RESIDENT_PREFETCH=$pf timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:lasso_resident_kernel -s 2 -c 2 --csv --log-file gpurun_out/r2ac_pf$pf.csv \
  python bench.py --gpus 1 --steps 20 --warmup 5 --legs fista --repeats 2 > gpurun_out/r2ac_ncu.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/r2ac_pf$pf.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; mi, vi, ui = h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
for r in rows[hdr + 1:]:
    if len(r) > vi: print('prefetch=$pf', r[mi], r[vi], r[ui])
PY
done
