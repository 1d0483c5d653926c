"""Per-source-line sample totals of an ncu source page: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv; python tools/ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
        d = dict(zip(hdr, r))
        out.append((int(d['# Samples'] or 0), int(d['Instructions Executed'] or 0), cur, int(r[0]), r[1].strip()[:120]))
out.sort(reverse=True)
tot = sum(o[0] for o in out)
print('total samples', tot)
for s, i, f, l, src in out[:n]:
    print(f'{s:7d} {100 * s / max(tot, 1):5.1f}% inst={i:10d} {f}:{l}  {src}')
