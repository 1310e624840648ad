#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export and a launch-list csv into a small text report (for profiles/)."""
import csv
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return None


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:78s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        st = [(hdr[i], num(r[i])) for i in range(len(hdr))
              if hdr[i].startswith('smsp__average_warps_issue_stalled_') and hdr[i].endswith('_per_issue_active.ratio')]
        st = sorted([x for x in st if x[1] is not None], key=lambda x: -x[1])[:6]
        print('  top stall reasons (warps per issue-active cycle):')
        for k, v in st:
            print(f"    {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
        rd, wr = num(r[hdr.index('dram__bytes_read.sum')]), num(r[hdr.index('dram__bytes_write.sum')])
        ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tot = rd * scale[ur] + wr * scale[uw]
        print(f"  dram traffic per launch: {tot / 1e9:.4f} GB")
        break  # one launch is enough


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(num(r[vi]))
    tot = sum(sum(v) for v in d.values())
    print('launch list (gpu__time_duration.sum, ns; cold-cache serialised -- compare shares):')
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {len(v):4d} x {sum(v) / len(v):12.0f} ns  {100 * sum(v) / tot:5.1f}%  {k[:90]}")


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print('==', p)
        (launches if 'launches' in p else raw)(p)
