#!/usr/bin/env python
"""Stall samples of one launch of an .ncu-rep attributed to the warp ROLES of a warp-specialised kernel.
SASS instructions (address order) are assigned to the role whose source-line range they fall into; instructions inlined
from helper headers inherit the role of the nearest preceding instruction that maps to the kernel file.
    python benchmarks/ncu_roles.py rep --launch 0 --file conv_halo.cuh --roles producer:150-268,mma:269-325,loader:326-352,epilogue:353-466
"""
import argparse
import collections
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--launch', type=int, default=0)
    ap.add_argument('--file', required=True)
    ap.add_argument('--roles', required=True)
    ap.add_argument('--top', type=int, default=8)
    a = ap.parse_args()
    roles = []
    for part in a.roles.split(','):
        name, rng = part.split(':')
        lo, hi = rng.split('-')
        roles.append((name, int(lo), int(hi)))
    out = subprocess.run(['ncu', '-i', a.rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-skip', str(a.launch),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    inst = {}
    hdr = fname = cur = func = None
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
            stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            continue
        if hdr is None:
            continue
        if r[0] != '':
            cur = (fname, int(r[0]), r[1].strip())
            continue
        if not r[2].startswith('0x'):
            continue
        addr = int(r[2], 16)
        s = int(r[i_s]) if r[i_s].isdigit() else 0
        st = {h: int(r[i]) for i, h in stall_cols if r[i].isdigit() and int(r[i])}
        # an instruction can be listed under several lines (inlining): keep the entry from the kernel file if any
        if addr not in inst or cur[0] == a.file:
            inst[addr] = (cur, r[3].strip(), s, st)
    role = 'prologue'
    per_role = collections.OrderedDict()
    for addr in sorted(inst):
        (f, ln, src), sass, s, st = inst[addr]
        if f == a.file:
            for name, lo, hi in roles:
                if lo <= ln <= hi:
                    role = name
                    break
        d = per_role.setdefault(role, dict(samples=0, stalls=collections.Counter(), lines=collections.Counter(), insts=0))
        d['insts'] += 1
        d['samples'] += s
        for k, v in st.items():
            d['stalls'][k] += v
        if s:
            d['lines'][f'{f}:{ln} {sass.split()[0] if not sass.startswith("@") else sass.split()[1]}'] += s
    total = sum(d['samples'] for d in per_role.values()) or 1
    print(func)
    for name, d in per_role.items():
        print(f"{name:10s} {100.0 * d['samples'] / total:5.1f}% of samples, {d['insts']} SASS instructions; stalls: "
              + ', '.join(f'{k} {100.0 * v / max(d["samples"], 1):.0f}%' for k, v in d['stalls'].most_common(5)))
        for k, v in d['lines'].most_common(a.top):
            print(f'      {100.0 * v / max(d["samples"], 1):5.1f}%  {k}')


if __name__ == '__main__':
    main()
