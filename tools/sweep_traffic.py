"""DRAM bytes of one NMF sweep from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log of tools/prof_nmf.py: sums the kernels between the last two normalize_rows launches.
usage: python tools/sweep_traffic.py log.csv"""
import csv
import json
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ii, ki, mi, vi, ui = h.index('ID'), h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
    launches = {}
    order = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        lid = int(r[ii])
        if lid not in launches:
            launches[lid] = {'name': r[ki]}
            order.append(lid)
        v = float(r[vi].replace(',', ''))
        unit = r[ui]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0,
                 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1.0)
        launches[lid][r[mi]] = v * scale
    norm = [i for i, lid in enumerate(order) if 'normalize' in launches[lid]['name']]
    lo, hi = norm[-2] + 1, norm[-1] + 1
    sweep = [launches[lid] for lid in order[lo:hi]]
    rd = sum(k.get('dram__bytes_read.sum', 0.0) for k in sweep)
    wr = sum(k.get('dram__bytes_write.sum', 0.0) for k in sweep)
    ns = sum(k.get('gpu__time_duration.sum', 0.0) for k in sweep)
    out = {'dram_bytes_read': rd, 'dram_bytes_written': wr, 'dram_bytes': rd + wr, 'kernel_time_ms': ns / 1e6,
           'kernels': [{'name': k['name'][:80], 'ms': k.get('gpu__time_duration.sum', 0.0) / 1e6,
                        'read': k.get('dram__bytes_read.sum', 0.0), 'written': k.get('dram__bytes_write.sum', 0.0)}
                       for k in sweep if k.get('gpu__time_duration.sum', 0.0) > 1e5]}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
