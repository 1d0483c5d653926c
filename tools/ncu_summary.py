"""Summarise ncu outputs (run here, no GPU needed): launch list CSV -> per-kernel table; .ncu-rep -> key metrics."""
import collections
import csv
import subprocess
import sys


def launch_table(path, passes=1):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        n = r[ki].split('(')[0].replace('mpa::', '').replace('void ', '')
        v = float(r[vi].replace(',', '')) / 1e6
        seq.append((n, v))
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ['| kernel | launches | total ms | share |', '|---|---|---|---|']
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| `{n[:70]}` | {c} | {t:.3f} | {100 * t / tot:.1f} % |')
    out.append(f'| **all** | {sum(v[0] for v in agg.values())} | {tot:.3f} | |')
    return '\n'.join(out), seq


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__cycles_elapsed.avg.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_tensor.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second']


def rep_metrics(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    out = []
    names = [r[h.index('Kernel Name')] if 'Kernel Name' in h else '' for r in rows[2:]]
    out.append('| metric | unit | ' + ' | '.join(f'launch {i}' for i in range(len(rows) - 2)) + ' |')
    out.append('|---|---|' + '---|' * (len(rows) - 2))
    for k in KEYS:
        if k in h:
            i = h.index(k)
            out.append(f'| `{k}` | {rows[1][i]} | ' + ' | '.join(r[i] for r in rows[2:]) + ' |')
    return '\n'.join(out), names


if __name__ == '__main__':
    if sys.argv[1].endswith('.csv'):
        print(launch_table(sys.argv[1])[0])
    else:
        print(rep_metrics(sys.argv[1])[0])
