"""Per-step summary of an ncu launch list of a training bench (`--no-train-graph`): the launches between the last two LayerNorm forwards.

    python tools/step_summary.py gpurun_out/launches.csv [N top kernels] [--seq]"""
import collections
import re
import sys

sys.path.insert(0, __file__.rsplit('/', 1)[0])
from launch_summary import load

if __name__ == '__main__':
    seq = load(sys.argv[1])
    idx = [i for i, (k, v) in enumerate(seq) if 'layernorm_cf_kernel' in k or 'layernorm_cf_cp8_kernel' in k or 'layernorm_pix_cp8_kernel' in k]
    st = seq[idx[-2]:idx[-1]]
    print(f'{len(st)} launches per step, {sum(v for k, v in st):.0f} us')
    tot, cnt = collections.Counter(), collections.Counter()
    for k, v in st:
        k = re.sub(r'\(.*', '', k).replace('mpa::', '').replace('void ', '')
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
    for k, v in tot.most_common(top):
        print(f'{v / s * 100:5.1f}% {cnt[k]:4d} x {v / cnt[k]:7.1f}  {v:8.1f} {k[:70]}')
    if '--seq' in sys.argv:
        for i, (k, v) in enumerate(st):
            k = re.sub(r'\(.*', '', k).replace('mpa::', '').replace('void ', '')
            print(f'{i:4d} {v:8.1f} {k[:60]}')
