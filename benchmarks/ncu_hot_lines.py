#!/usr/bin/env python
"""Per source line stall samples of one launch of an .ncu-rep (read on the CPU box):
    python benchmarks/ncu_hot_lines.py gpurun_out/prof.ncu-rep --launch 0 [--top 25]
Groups the SASS rows of `ncu --page source --print-source cuda,sass --csv` under their CUDA source line and prints the
lines with the most warp-stall samples, the dominant stall reasons, and the opcodes that collected them."""
import argparse
import collections
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--launch', type=int, default=0)
    ap.add_argument('--top', type=int, default=25)
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-skip', str(a.launch),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, func, hdr = None, None, None
    lines = collections.OrderedDict()
    cur = None
    total = 0
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            func = r[1]
            continue
        if r[0] == 'Line No':
            hdr = r
            i_s = hdr.index('# Samples')
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            continue
        if hdr is None:
            continue
        if r[0] != '':
            cur = (fname, int(r[0]), r[1].strip())
            lines.setdefault(cur, dict(samples=0, stalls=collections.Counter(), ops=collections.Counter()))
            continue
        if cur is None:
            continue
        s = int(r[i_s]) if r[i_s].isdigit() else 0
        if s == 0:
            continue
        d = lines[cur]
        d['samples'] += s
        total += s
        op = r[3].split()[0] if r[3].split() else '?'
        if op.startswith('@'):
            op = r[3].split()[1]
        d['ops'][op] += s
        for i, h in stall_cols:
            v = int(r[i]) if r[i].isdigit() else 0
            if v:
                d['stalls'][h[6:]] += v
    print(func)
    print(f'total samples {total}')
    for (f, ln, src), d in sorted(lines.items(), key=lambda kv: -kv[1]['samples'])[:a.top]:
        st = ', '.join(f'{k} {v}' for k, v in d['stalls'].most_common(3))
        ops = ', '.join(f'{k} {v}' for k, v in d['ops'].most_common(3))
        print(f"{100.0 * d['samples'] / max(total, 1):5.1f}%  {f}:{ln:<4d} {src[:90]}\n        stalls: {st}   ops: {ops}")


if __name__ == '__main__':
    main()
