"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernels of ONE Adam step
(between two consecutive adam_kernel launches), grouped by kernel.  Usage: launch_summary.py CSV"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1.0)
        name = re.sub(r'^void\s+', '', r['Kernel Name'])
        name = re.sub(r'\(.*$', '', name)
        name = re.sub(r'^.*?(?:unnamed>|anonymous namespace\))::', '', name)
        name = re.sub(r'^la::', '', name)
        rows.append((name, us))
    adam = [i for i, (n, _) in enumerate(rows) if 'adam_kernel' in n]
    if len(adam) < 2:
        print('no complete step in the capture window'); return
    k = min(range(len(adam) - 1), key=lambda i: adam[i + 1] - adam[i])      # a window inside one la_augment call
    step = rows[adam[k] + 1:adam[k + 1] + 1]
    tot = sum(u for _, u in step)
    print('ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised)')
    print(f'one Adam step: {len(step)} kernels, {tot:.1f} us')
    agg = collections.OrderedDict()
    for n, u in step:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += u
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'  {u:8.1f} us {100 * u / tot:5.1f}% n={c:3d} {n}')


if __name__ == '__main__':
    main(sys.argv[1])
