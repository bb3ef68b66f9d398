"""Per-kernel resource table of the built library (registers, stack frame, static shared memory) from
``cuobjdump -res-usage``; with ``--ptxas FILE`` (the stderr of ``nvcc ... -Xptxas -v -c csrc/X.cu``) the spill bytes too.

    python tools/kernel_resources.py [--ptxas /tmp/ptxas_tapgemm.txt] > profiles/r2_kernel_resources.txt
"""
import argparse
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'latentaugment_b200', 'liblatentaugment_b200.so')


def demangle(sym):
    name = subprocess.run(['c++filt', sym], capture_output=True, text=True).stdout.strip()
    name = name.replace('(anonymous namespace)::', '')
    return re.sub(r'\(.*', '', name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ptxas', default='')
    a = ap.parse_args()
    txt = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    rows = []
    for m in re.finditer(r'Function (\S+):\n\s*(REG:\d+ .*)', txt):
        res = dict(kv.split(':') for kv in m.group(2).split())
        rows.append((demangle(m.group(1)), int(res['REG']), int(res.get('STACK', 0)), int(res.get('SHARED', 0)), int(res.get('LOCAL', 0))))
    rows.sort()
    print(f'{len(rows)} kernels in {os.path.basename(LIB)} (sm_100a cubins; cuobjdump -res-usage)')
    print(f"{'kernel':72s} {'regs':>5s} {'stack':>6s} {'static smem':>11s}")
    for r in rows:
        print(f'{r[0][:72]:72s} {r[1]:5d} {r[2]:6d} {r[3]:11d}')
    if a.ptxas:
        print('\nptxas -v: kernels with register spills (the tensor-core kernels are capped at 168 registers by their 384-thread launch bound)')
        t = open(a.ptxas).read()
        for m in re.finditer(r'Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', t):
            if int(m.group(3)) or int(m.group(4)):
                print(f'{demangle(m.group(1))[:72]:72s} stack {int(m.group(2)):4d} B, spill stores {int(m.group(3)):4d} B, spill loads {int(m.group(4)):4d} B')


if __name__ == '__main__':
    main()
