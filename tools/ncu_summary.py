"""Summarise an .ncu-rep: per-kernel key metrics (raw page) and, per kernel, the stall mix and the
hottest source lines (source page).  Usage: python tools/ncu_summary.py REPORT [--src KERNEL_INDEX]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__registers_per_thread', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h = r[0]
    ki = h.index('Kernel Name')
    for row in r[2:]:
        print('==', row[0], row[ki][:90])
        for k in KEYS:
            for i, c in enumerate(h):
                if c.endswith(k) or c == k:
                    print(f'   {k:85s} {row[i]} {r[1][i]}')
                    break


def source(rep, idx):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'] if False else
                         ['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name"')
    blk = '"Kernel Name"' + blocks[idx + 1]
    r = list(csv.reader(io.StringIO(blk)))
    hi = [i for i, row in enumerate(r) if row and row[0] == 'Address'][0]
    h = r[hi]
    rows = [row for row in r[hi + 1:] if len(row) == len(h) and row[0] != 'Address']
    stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    tot = collections.Counter()
    for row in rows:
        for c in stalls:
            try:
                tot[c] += float(row[h.index(c)])
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    print(r[0][1][:100])
    print('stalls:', ', '.join(f'{k[6:]} {100 * v / s:.0f}%' for k, v in tot.most_common(8)))
    ie, si, src = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
    total = sum(float(x[ie] or 0) for x in rows)
    print('warp instructions', total)
    op = collections.Counter()
    for x in rows:
        m = x[src].split()
        name = (m[1] if m[0].startswith('@') else m[0]).split('.')[0]
        op[name] += float(x[ie] or 0)
    print('mix:', ', '.join(f'{k} {100 * v / total:.1f}%' for k, v in op.most_common(16)))
    top = sorted(rows, key=lambda x: -float(x[si] or 0))[:25]
    for t in top:
        big = {c[6:]: t[h.index(c)] for c in stalls if float(t[h.index(c)] or 0) > 0.25 * float(t[si] or 1)}
        print(f'  {t[si]:>6s} {t[src][:90]:90s} {big}')


if __name__ == '__main__':
    rep = sys.argv[1]
    if '--src' in sys.argv:
        source(rep, int(sys.argv[sys.argv.index('--src') + 1]))
    else:
        raw(rep)
