"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel; optionally print the last N launches in order.

    python tools/launch_summary.py profiles/r01_launches_….csv [N]"""
import collections
import csv
import sys


def load(fn):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
        seq.append((r[ki][:80], v))
    return seq


if __name__ == '__main__':
    seq = load(sys.argv[1])
    tot, cnt = collections.Counter(), collections.Counter()
    for k, v in seq:
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    print(f'{len(seq)} launches, {s:.0f} us')
    for k, v in tot.most_common(25):
        print(f'  {v / s * 100:5.1f}% {cnt[k]:5d} x {v / cnt[k]:8.1f} us  {k}')
    if len(sys.argv) > 2:
        for k, v in seq[-int(sys.argv[2]):]:
            print(f'{v:9.1f}  {k}')
