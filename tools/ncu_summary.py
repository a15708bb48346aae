"""Summarise an exported ncu report: raw-page key metrics, stall breakdown, and per-region samples of the source page.
usage: python tools/ncu_summary.py raw.csv [src.csv]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'sm__cycles_elapsed.avg']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('%-70s %s %s' % (w, r[i][:110], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if re.match(r'smsp__pcsamp_warps_issue_stalled_', h) and not h.endswith('not_issued'):
            try:
                st.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print('stalls: ' + ', '.join('%s %.1f%%' % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:8]))
    print('---')
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    iS, iN = hdr.index('Source'), hdr.index('# Samples')
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) <= iN or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            data.append((r[iS], int(r[iN])))
        except ValueError:
            pass
    tot = sum(d[1] for d in data) or 1
    idx = [i for i, d in enumerate(data) if 'DMMA' in d[0]]
    pre = sum(d[1] for d in data[:idx[0]])
    main = sum(d[1] for d in data[idx[0]:idx[-1] + 1])
    post = sum(d[1] for d in data[idx[-1] + 1:])
    print('instructions %d (DMMA %d); samples: before mainloop %.1f%%, mainloop %.1f%%, epilogue %.1f%%'
          % (len(data), len(idx), 100 * pre / tot, 100 * main / tot, 100 * post / tot))
    op = collections.Counter()
    for s_, n in data:
        m = re.sub(r'^@!?U?P\d+\s+', '', s_.strip()).split()[0] if s_.strip() else '?'
        op[m] += n
    print('by opcode: ' + ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in op.most_common(12)))
    for d in sorted(data, key=lambda d: -d[1])[:12]:
        print('  %5.1f%%  %s' % (100 * d[1] / tot, d[0][:100]))
