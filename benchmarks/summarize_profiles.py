#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small tracked summaries under profiles/.

    python benchmarks/summarize_profiles.py launches gpurun_out/launches_r1_final.csv profiles/r1_launches_by_kernel.csv
    python benchmarks/summarize_profiles.py full     gpurun_out/prof_r1_final.ncu-rep profiles/r1_ncu_full_summary.csv
    python benchmarks/summarize_profiles.py conv     gpurun_out/conv_launches_r1.csv profiles/r1_conv_launches.csv profiles/r1_conv_traffic.json

`launches`: the `ncu --metrics gpu__time_duration.sum --csv` launch list of one training iteration -> per kernel name:
launches, total ms, share of the summed kernel time (cold-cache, serialised: compare SHARES, not absolutes).
`conv`: the multi-metric launch list of the convolution kernels (time, DRAM bytes, tensor-pipe activity per launch) ->
one row per launch plus the per-family averages `bench.py` reports as `roofline.traffic`.
`full`: an `ncu --set full` report -> one row per profiled launch with the metrics the roofline discussion uses.
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def short(name):
    name = name.replace('void ', '').replace('sgb::', '')
    i = name.find('(')
    return name[:i] if i > 0 else name


def launches(src, dst):
    lines = open(src).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = short(r['Kernel Name'])
        if k.startswith('at::') or 'cutlass' in k or 'cublas' in k.lower():
            k = '[torch/aten] ' + k[:70]
        agg[k][0] += 1
        agg[k][1] += float(r['Metric Value']) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_ms', 'share_of_kernel_time'])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f'{v[1]:.4f}', f'{v[1] / tot:.4f}'])
        w.writerow(['TOTAL', len(rows), f'{tot:.4f}', '1.0'])
    print(f'{len(rows)} launches, {tot:.2f} ms of kernel time -> {dst}')


def full(src, dst):
    # src: an .ncu-rep, or the `ncu -i rep --page raw --csv` export made on the GPU box (the reports themselves are too large
    # to bring back: gpurun_out/ is capped at 64 MiB)
    if src.endswith('.csv'):
        out = open(src).read()
    else:
        out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in FULL_METRICS if m in hdr]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([f'{m} [{units[i]}]' if units[i] else m for m, i in idx])
        for r in rows[2:]:
            w.writerow([short(r[i]) if m == 'Kernel Name' else r[i] for m, i in idx])
    print(f'{len(rows) - 2} profiled launches -> {dst}')


def conv(src, dst, dst_json):
    import json
    lines = open(src).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        d = per.setdefault(r['ID'], {'kernel': short(r['Kernel Name']), 'grid': r['Grid Size']})
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'grid', 'ms', 'dram_bytes', 'tensor_pipe_active_pct'])
        for d in per.values():
            ms = d['gpu__time_duration.sum'] / 1e6
            nbytes = d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']
            tp = d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
            w.writerow([d['kernel'], d['grid'], f'{ms:.4f}', int(nbytes), f'{tp:.2f}'])
            k = 'conv_wgrad_tc' if 'wgrad' in d['kernel'] else ('conv_fwd_small' if 'small' in d['kernel'] else 'conv_fwd_tc')
            fam[k][0] += 1; fam[k][1] += ms; fam[k][2] += nbytes; fam[k][3] += tp * ms
    out = {k: {'launches': v[0], 'total_ms': round(v[1], 3), 'dram_bytes_per_launch': v[2] / v[0],
               'time_weighted_tensor_pipe_active_pct': round(v[3] / v[1], 2)} for k, v in fam.items()}
    out['source'] = ('ncu --nvtx --nvtx-include sgb_timed -k regex:conv... --metrics dram__bytes_read.sum,dram__bytes_write.sum,'
                     'gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active on `bench.py --no-graphs '
                     '--steps 1 --warmup 3` (one iteration with all four phases, ffhq256 batch 32, TF32); ' + src)
    with open(dst_json, 'w') as f:
        json.dump(out, f, indent=1)
    print(f'{len(per)} launches -> {dst}, {dst_json}')


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'conv': conv}[sys.argv[1]](*sys.argv[2:])
