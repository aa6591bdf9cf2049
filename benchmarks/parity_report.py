#!/usr/bin/env python
"""Parity table: the reference's unchanged callers (train_parts G / D / SG2Loss / R1reg / PPLreg from baseline/_ref) against the
reference-generated golden gradients (tests/golden/net_tiny.npz, CPU fp32 impl='ref'), per arithmetic mode and per training
phase, for BOTH backends:

    sgb200     the libsgb200 kernels behind sgb200.install()
    reference  the reference's own CUDA path on the same GPU (cuDNN convolutions with the same allow_tf32, its plugins)

The second column calibrates the first: what the reference's own GPU arithmetic (cuDNN TF32 / fp16) does to the same
gradients under the same metric.  Metric: per-tensor max-norm relative error with the 1 % floor of tests/helpers.py.

    python benchmarks/parity_report.py [--md profiles/r2_parity_report.md]       (spawns one process per backend)
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

MODES = {
    'strict': dict(tf32=False, num_fp16_res=0, conv_clamp=None),
    'tf32': dict(tf32=True, num_fp16_res=0, conv_clamp=None),
    'fp16': dict(tf32=True, num_fp16_res=2, conv_clamp=256),
}


def worker(backend):
    import numpy as np
    import torch
    from benchmarks import ref_harness as H
    H.import_reference(backend)
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'net_tiny.npz'))
    meta = json.loads(str(z['meta']))
    c = meta['cfg']
    out = {}
    for mode, m in MODES.items():
        torch.backends.cudnn.allow_tf32 = m['tf32']
        torch.backends.cuda.matmul.allow_tf32 = False
        w = dict(res=c['img_resolution'], batch_gpu=meta['n'], z_dim=c['z_dim'], w_dim=c['w_dim'], map_layers=c['map_layers'],
                 channel_base=c['channel_base'], d_arch=c['d_arch'], mbstd=c['mbstd_group_size'], r1_gamma=meta['r1_gamma'], ppl=True,
                 style_mixing_prob=0.0, num_fp16_res=m['num_fp16_res'], conv_clamp=m['conv_clamp'], ema_kimg=10.0, g_attn=(), d_attn=())
        tr = H.RefCallerTrainer(w, 'cuda', backend, noise_mode='const', channel_max=c['channel_max'], use_ema=False)
        tr.G.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('G.')})
        tr.D.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('D.')})
        zz = torch.from_numpy(z['z']).cuda()
        real = torch.from_numpy(z['real']).cuda()
        pl_noise = torch.from_numpy(z['pl_noise'])
        orig = torch.randn_like
        res = {}
        for phase, tag in [('Gmain', 'G.'), ('Dmain', 'D.'), ('Dreg', 'D.'), ('Greg', 'G.')]:
            torch.randn_like = lambda t, **k: pl_noise.to(device=t.device, dtype=t.dtype)
            try:
                ph = tr.phase_grads(phase, real, zz, meta['gains'][phase])
            finally:
                torch.randn_like = orig
            keys = [k for k in z.files if k.startswith(f'{phase}.grad.{tag}')]
            floor = 1e-2 * max(float(np.abs(z[k]).max()) for k in keys)
            named = dict(ph.module.named_parameters())
            errs = []
            for k in keys:
                name = k[len(f'{phase}.grad.{tag}'):]
                ref = torch.from_numpy(z[k]).double()
                got = named[name].grad.detach().double().cpu()
                errs.append(((got - ref).abs().max().item() / max(ref.abs().max().item(), floor), name))
            errs.sort(reverse=True)
            res[phase] = dict(worst=errs[0][0], tensor=errs[0][1], median=errs[len(errs) // 2][0], n_over_1e2=sum(e > 1e-2 for e, _ in errs),
                              n=len(errs))
        out[mode] = res
    out_a = config_a(H, backend)
    print('PARITY_JSON ' + json.dumps(dict(backend=backend, result=out, config_a=out_a,
                                            plugins=H.plugin_status() if backend == 'reference' else None)), flush=True)


def config_a(H, backend):
    """BASELINE configs[0] at its real size (sg2ada.yaml: 64x64, batch 8, 512 channels, D 'orig'): gradients of Gmain / Dmain /
    Dreg from the unchanged callers on the GPU against the SAME reference modules run on the CPU in fp32 (impl='ref').  The
    'reference' worker runs first, computes that CPU golden and leaves it in the system temp directory for the 'sgb200' worker
    (whose ops refuse CPU tensors).  ~600 MB in fp64: it must NOT sit in gpurun_out/, whose size is capped at 64 MiB."""
    import numpy as np
    import torch
    import tempfile
    path = os.path.join(tempfile.gettempdir(), 'sgb200_parity_config_a_golden.npz')
    g = torch.Generator().manual_seed(5)
    zz = torch.randn(8, 512, generator=g)
    real = torch.rand(8, 3, 64, 64, generator=g) * 2 - 1

    def build(dev):
        tr = H.RefCallerTrainer('sg2ada64', dev, backend, noise_mode='const', use_ema=False, seed=3)
        with torch.no_grad():
            for n_, p in tr.G.named_parameters():
                if n_.endswith('noise_strength'):
                    p.fill_(0.1)
        return tr

    def grads(tr, dev):
        res = {}
        for phase, gain in (('Gmain', 1), ('Dmain', 1), ('Dreg', 4)):
            ph = tr.phase_grads(phase, real.to(dev), zz.to(dev), gain)
            res[phase] = {k: p.grad.detach().double().cpu().numpy() for k, p in ph.module.named_parameters() if p.grad is not None}
        return res

    if backend == 'reference':
        torch.backends.cudnn.allow_tf32 = False
        gold = grads(build('cpu'), 'cpu')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, **{f'{ph}/{k}': v for ph, d in gold.items() for k, v in d.items()})
    if not os.path.exists(path):
        return None
    z = np.load(path)
    out = {}
    for mode in ('strict', 'tf32'):
        torch.backends.cudnn.allow_tf32 = mode == 'tf32'
        got = grads(build('cuda'), 'cuda')
        out[mode] = {}
        for phase, d in got.items():
            keys = [k for k in z.files if k.startswith(phase + '/')]
            floor = 1e-2 * max(float(np.abs(z[k]).max()) for k in keys)
            errs = sorted((float(np.abs(d[k.split('/', 1)[1]] - z[k]).max()) / max(float(np.abs(z[k]).max()), floor), k.split('/', 1)[1]) for k in keys)
            out[mode][phase] = dict(worst=errs[-1][0], tensor=errs[-1][1], median=errs[len(errs) // 2][0])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--worker', default=None)
    ap.add_argument('--md', default=None)
    a = ap.parse_args()
    if a.worker:
        worker(a.worker)
        return
    rows = {}
    for backend in ('reference', 'sgb200'):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', backend], capture_output=True, text=True, timeout=1500)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('PARITY_JSON ')]
        if not line:
            print(f'# {backend}: failed\n{r.stderr[-2000:]}')
            continue
        rows[backend] = json.loads(line[-1][len('PARITY_JSON '):])
        print(json.dumps(rows[backend]))
    lines = ['| mode | phase | sgb200 worst (tensor) | sgb200 median | reference GPU path worst (tensor) | reference median |', '|---|---|---|---|---|---|']
    for mode in MODES:
        for phase in ('Gmain', 'Dmain', 'Dreg', 'Greg'):
            cells = []
            for backend in ('sgb200', 'reference'):
                r = rows.get(backend, {}).get('result', {}).get(mode, {}).get(phase)
                cells += [f"{r['worst']:.2e} ({r['tensor']})", f"{r['median']:.2e}"] if r else ['n/a', 'n/a']
            lines.append(f'| {mode} | {phase} | ' + ' | '.join(cells) + ' |')
    lines += ['', 'Config A at its real size (sg2ada.yaml 64x64, batch 8, 512 channels), golden = the same reference modules on the CPU (fp32):', '',
              '| mode | phase | sgb200 worst (tensor) | sgb200 median | reference GPU path worst (tensor) | reference median |', '|---|---|---|---|---|---|']
    for mode in ('strict', 'tf32'):
        for phase in ('Gmain', 'Dmain', 'Dreg'):
            cells = []
            for backend in ('sgb200', 'reference'):
                r = ((rows.get(backend, {}).get('config_a') or {}).get(mode) or {}).get(phase)
                cells += [f"{r['worst']:.2e} ({r['tensor']})", f"{r['median']:.2e}"] if r else ['n/a', 'n/a']
            lines.append(f'| {mode} | {phase} | ' + ' | '.join(cells) + ' |')
    print('\n'.join(lines))
    if a.md:
        with open(a.md, 'w') as fh:
            fh.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
