#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python profiles/summarize.py launches <launches.csv>          per-kernel totals / shares of the step
  python profiles/summarize.py full <prof.ncu-rep>              key metrics of an `ncu --set full` capture
  python profiles/summarize.py stalls <prof.ncu-rep>            warp-stall cycles per issued instruction, top reasons per kernel
  python profiles/summarize.py metrics <prof.ncu-rep> <workload> <n_gpus> <git-hash> [fp64_peak_tflops]
                                                                 -> profiles/ncu_metrics.json, read by bench.py for
                                                                    roofline.traffic / roofline.fp64 (per-launch figures)
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


def _raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return r[0], r[1], r[2:]


def _num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}


def stalls(path):
    hdr, units, rows = _raw(path)
    cols = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
    extra = ['smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
             'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
    print('# warp-stall cycles per issued instruction (ncu --set full), top reasons per kernel')
    for row in rows:
        print(row[hdr.index('Kernel Name')][:60])
        top = sorted(((_num(row[i]) or 0.0, hdr[i].split('stalled_')[1].split('_per_issue')[0]) for i in cols), reverse=True)[:6]
        for v, n in top:
            print(f'    {v:6.2f}  {n}')
        for e in extra:
            if e in hdr:
                print(f'    {e} = {row[hdr.index(e)]}')


def metrics(path, workload, n_gpus, git, fp64_peak='37.2'):
    """Per-launch figures of the tendency kernels for bench.py: DRAM bytes (read + write), FP64 flops (DFMA = 2, DMUL / DADD = 1,
    SASS op counters), FP64 pipe utilisation; both momentum instances (full-order + masked tiles) are summed per launch pair."""
    import json
    import os
    hdr, units, rows = _raw(path)

    def val(row, name):
        if name not in hdr:
            return None
        i = hdr.index(name)
        v = _num(row[i])
        return None if v is None else v * _SCALE.get(units[i], 1.0)

    agg = {}
    for row in rows:
        name = row[hdr.index('Kernel Name')]
        key = 'momentum' if 'k_momentum_tiled' in name else 'tracers' if 'k_tracers_tiled' in name else None
        if key is None:
            continue
        gen = ', 1, 1>' in name.replace('(bool)', '').replace('(int)', '') or '<16, 16, 1,' in name
        a = agg.setdefault(key, {'dram_bytes': 0.0, 'fp64_flops': 0.0, 'ms': 0.0, 'launches': 0, 'fp64_pipe_frac': None})
        a['dram_bytes'] += (val(row, 'dram__bytes_read.sum') or 0.0) + (val(row, 'dram__bytes_write.sum') or 0.0)
        # thread-level op counts: the .sum counter when the capture has it, else (sum over SMSPs per cycle) x elapsed cycles
        cyc = val(row, 'smsp__cycles_elapsed.avg') or 0.0
        f = [val(row, f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum')
             or (val(row, f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed') or 0.0) * cyc
             for op in ('dfma', 'dmul', 'dadd')]
        a['fp64_flops'] += 2 * f[0] + f[1] + f[2]
        a['ms'] += val(row, 'gpu__time_duration.sum') or 0.0
        a['launches'] += 1
        if not gen:
            pf = val(row, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')
            a['fp64_pipe_frac'] = None if pf is None else pf / 100.0
    out = {'workload': workload, 'n_gpus': int(n_gpus), 'git': git, 'source': os.path.basename(path) + ' (ncu --set full --clock-control none)',
           'fp64_peak_tflops': float(fp64_peak), 'fp64_peak_source': 'tools/fp64_peak.cu on this pool\'s B200 (= 64 DFMA/clk/SM x 148 SM x 2 x 1.965 GHz)',
           'note': 'per step: the two launches (full-order + masked tiles) of each tendency kernel summed; ncu times are cold-cache and serialised'}
    out.update(agg)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_metrics.json')
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    fn = {'launches': launches, 'full': full, 'stalls': stalls, 'metrics': metrics}[sys.argv[1]]
    fn(*sys.argv[2:])
