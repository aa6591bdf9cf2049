#!/usr/bin/env python
"""Per-kernel SASS instruction summary of the shipped objects (style-big-gan_b200/sgb200/lib/*.o) -> profiles/sass_summary.txt

    python benchmarks/sass_summary.py [out.txt]

Counts, per kernel, the mnemonics that show which hardware path it uses (B200_PROFILING.md): UTC*MMA (tcgen05.mma; the
`.2CTA` suffix = cta_group::2), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UTMALDG/UTMASTG (TMA tensor loads / stores),
UBLKCP (cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier), HMMA/IMMA (mma.sync -- none expected), RED/ATOM, plus the
instruction count.  Runs without a GPU (cuobjdump only).
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'style-big-gan_b200', 'sgb200', 'lib')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', '2CTA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS',
        'HMMA', 'IMMA', 'RED', 'ATOM', 'LDG', 'STG', 'LDS', 'STS', 'FFMA']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for m, d in zip(names, out):
        d = d.replace('void ', '').replace('sgb::', '')
        i = d.find('(')
        res[m] = d[:i] if i > 0 else d
    return res


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'sass_summary.txt')
    rows = []
    for obj in sorted(glob.glob(os.path.join(LIB, '*.o'))):
        sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        fn, cnt, total = None, None, 0
        per = collections.OrderedDict()
        for line in sass.splitlines():
            m = re.search(r'Function : (\S+)', line)
            if m:
                fn = m.group(1)
                per[fn] = [collections.Counter(), 0]
                continue
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m and fn:
                op = m.group(1)
                per[fn][1] += 1
                base = op.split('.')[0]
                if base in KEYS:
                    per[fn][0][base] += 1
                if base.startswith('UTC') and base.endswith('MMA') and '2CTA' in op:
                    per[fn][0]['2CTA'] += 1
        names = demangle(list(per))
        for fn, (c, n) in per.items():
            rows.append((os.path.basename(obj), names[fn], n, c))
    with open(dst, 'w') as f:
        f.write('# SASS summary of style-big-gan_b200/sgb200/lib/*.o (cuobjdump -sass, sm_100a); produced by benchmarks/sass_summary.py\n')
        f.write('# columns: object | kernel | instructions | non-zero counts of ' + ' '.join(KEYS) + '\n')
        for obj, name, n, c in rows:
            f.write(f'{obj} | {name} | {n} | ' + ' '.join(f'{k}={c[k]}' for k in KEYS if c[k]) + '\n')
        tot = collections.Counter()
        for _, _, _, c in rows:
            tot.update(c)
        f.write('# totals: ' + ' '.join(f'{k}={tot[k]}' for k in KEYS) + f' kernels={len(rows)}\n')
    print(f'{len(rows)} kernels -> {dst}')


if __name__ == '__main__':
    main()
