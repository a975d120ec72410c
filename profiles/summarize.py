#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python profiles/summarize.py launches <launches.csv>          per-kernel totals / shares of the step
  python profiles/summarize.py full <prof.ncu-rep>              key metrics of an `ncu --set full` capture
"""
import collections
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}[row['Metric Unit']]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time over the captured launches (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} n={v[0]:4d} total={v[1]:9.3f} ms  avg={v[1] / v[0]:8.3f} ms  share={100 * v[1] / tot:5.1f}%")


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr = r[0]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    for row in r[2:]:
        print(row[hdr.index('Kernel Name')][:60])
        for i in idx[1:]:
            print(f"    {hdr[i]:70s} {row[i]:>16s} {r[1][i]}")


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
