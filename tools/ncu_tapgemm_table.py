"""Per-launch table of an `ncu --metrics ... --csv` capture of the 26 tap-GEMM launches of one Adam step (13 forward, then 13
data-gradient launches top layer first): duration, tensor-pipe activity, DRAM traffic, IPC, shared-memory bank conflicts.
Usage: python tools/ncu_tapgemm_table.py CAPTURE.csv [res]"""
import collections
import csv
import io
import re
import sys


def main():
    txt = open(sys.argv[1]).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    by = collections.OrderedDict()
    for r in rows:
        try:
            val = float(r['Metric Value'].replace(',', ''))
        except ValueError:
            val = float('nan')
        by.setdefault(r['ID'], {'name': r['Kernel Name'], 'grid': r['Grid Size']})[r['Metric Name']] = val
    names = []
    r = 4
    while r <= res:
        if r > 4:
            names.append(f'b{r}.conv0')
        names.append(f'b{r}.conv1')
        r *= 2
    L = list(by.values())
    n = len(names)
    tot_t = tot_d = 0.0
    print(f'{"launch":18s} {"kernel<BN,EPI,pair>":>20s} {"grid":>6s} {"us":>8s} {"tensor pipe %":>14s} {"DRAM MB":>9s} {"IPC":>5s} {"smem bank conflicts (M)":>24s}')
    for i, v in enumerate(L):
        m = re.search(r'tapgemm_kernel<(\d+), (\d+), (\d+)>', v['name'])
        lay = names[i] + ' fwd' if i < n else names[2 * n - 1 - i] + ' dgrad'
        t = v['gpu__time_duration.sum'] / 1e3
        d = (v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / 1e6
        tot_t += t
        tot_d += d
        grid = re.sub(r'[(), ]+', ' ', v['grid']).split()[0]
        print(f"{lay:18s} {'<' + ','.join(m.groups()) + '>':>20s} {grid:>6s} {t:8.1f} "
              f"{v['sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']:14.1f} {d:9.1f} "
              f"{v['sm__inst_executed.avg.per_cycle_elapsed']:5.2f} {v['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] / 1e6:24.2f}")
    print(f'sum: {tot_t:.1f} us, DRAM read + write {tot_d / 1e3:.3f} GB   (ncu: cold cache, serialised; EPI 1 forward, 2 backward, 4 bf16 store)')
    if len(sys.argv) > 3:
        import json
        json.dump({'dram_bytes_read_plus_write': tot_d * 1e6, 'launches': len(L), 'source': sys.argv[1],
                   'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the tap-GEMM launches of one Adam step'}, open(sys.argv[3], 'w'))


if __name__ == '__main__':
    main()
