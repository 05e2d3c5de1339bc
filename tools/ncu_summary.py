#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/rNN_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
           'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_active']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    tot, n = collections.defaultdict(lambda: [0.0, 0]), 0
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        v = {'ns': v / 1e3, 'ms': v * 1e3, 's': v * 1e6}.get(row['Metric Unit'], v)
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('uocr::', '')
        tot[name][0] += v
        tot[name][1] += 1
        n += 1
    total = sum(v[0] for v in tot.values())
    with open(dst, 'w') as out:
        out.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches):\n'
                  f'# compare SHARES, not absolutes.  {n} launches, {total:.1f} us in total\n')
        out.write(f'{"us":>10s} {"calls":>6s} {"share":>7s}  kernel\n')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            out.write(f'{v[0]:10.1f} {v[1]:6d} {100 * v[0] / total:6.1f}%  {k}\n')
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    with open(dst, 'w') as out:
        out.write(f'# ncu --set full --clock-control none, from {src}\n')
        for col in ('Kernel Name', 'Grid Size', 'Block Size'):
            if col in head:
                i = head.index(col)
                out.write(f'{col}: {[r[i][:100] for r in rows[2:]]}\n')
        for m in METRICS:
            if m in head:
                i = head.index(m)
                out.write(f'{m} [{units[i]}]: {[r[i] for r in rows[2:]]}\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
