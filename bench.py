#!/usr/bin/env python
"""bench.py -- StyleGAN2 G+D training throughput on the sgb200 ops (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ffhq256|f1024|sg2ada64] [--impl ours|reference]

One "step" = one reference training iteration at the workload's per-GPU batch: Gmain + Dmain every step, Dreg
(R1) every d_reg_interval steps, Greg (path length) every g_reg_interval steps, Adam updates, G_ema
(train_parts/trainers.py:699-761).  The default K = 16 covers exactly one period of the lazy-regularisation
schedule (1 Greg + 4 Dreg), so ms_per_step is the reference's amortised iteration cost.  Warm-up runs every
phase at least once.

Prints ONE JSON line (rank 0).  `value` = img/s with the real-image batch already resident in HBM; `e2e` = the
same loop fed from pinned host memory (H2D copy of the uint8 batch + D2H read of the losses inside the timed
region).  `roofline` is for the kernel family with the largest share of device time, measured with CUDA events
around its launches during the timed loop; `cpu_baseline` is the CPU oracle (oracle/ref_networks.py = the
reference's impl='ref' arithmetic) on this box's host cores, on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tc_burst=p['bf16_tflops'], tc_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def workload_config(name, **over):
    from sgb200 import training
    return {'ffhq256': training.config_ffhq256, 'f1024': training.config_f1024, 'sg2ada64': training.config_sg2ada64}[name](**over)


WORKLOAD_DESC = {
    'ffhq256': 'ffhq_sg2.yaml StyleGAN2 256x256 batch 32/GPU G+D train step (fp32, R1 every 4, PPL every 16)',
    'f1024': 'StyleGAN2 config-f 1024x1024 batch 4/GPU G+D train step (fp16 top-4 res, clamp 256, R1 every 4, PPL every 16)',
    'sg2ada64': 'sg2ada.yaml StyleGAN2 64x64 batch 8 G+D train step (fp32, R1 every 4)',
}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's impl='ref' path, bounded sample of the same workload
def cpu_iteration_rate(cfg, batch, reps=1):
    """img/s of one amortised training iteration (Gmain + Dmain + Greg/gi + Dreg/di) on the host CPU at
    per-step batch `batch` (Greg uses batch // 2 like the reference)."""
    from oracle import ref_networks as RN
    ocfg = RN.NetConfig(img_resolution=cfg.img_resolution, z_dim=cfg.z_dim, w_dim=cfg.w_dim, channel_base=cfg.channel_base,
                        channel_max=cfg.channel_max, map_layers=cfg.map_layers, num_fp16_res=0, conv_clamp=cfg.conv_clamp,
                        d_arch=cfg.d_arch, mbstd_group_size=min(cfg.mbstd_group_size, batch))
    GP = RN.init_params(RN.g_param_shapes(ocfg), seed=0)
    DP = RN.init_params(RN.d_param_shapes(ocfg), seed=1)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(batch, cfg.z_dim, generator=g)
    real = torch.rand(batch, 3, cfg.img_resolution, cfg.img_resolution, generator=g) * 2 - 1
    noise = 'const'
    t = {}
    for _ in range(reps):
        t0 = time.perf_counter(); RN.phase_gmain(GP, DP, z, ocfg, ocfg, noise=noise); t['Gmain'] = time.perf_counter() - t0
        t0 = time.perf_counter(); RN.phase_dmain(GP, DP, z, real, ocfg, ocfg, noise=noise); t['Dmain'] = time.perf_counter() - t0
        t['Dreg'] = t['Greg'] = 0.0
        if cfg.use_r1:
            t0 = time.perf_counter(); RN.phase_dreg(DP, real, ocfg, r1_gamma=cfg.r1_gamma); t['Dreg'] = time.perf_counter() - t0
        if cfg.use_ppl:
            nb = max(batch // cfg.pl_batch_shrink, 1)
            pl = torch.randn(nb, 3, cfg.img_resolution, cfg.img_resolution, generator=g)
            t0 = time.perf_counter(); RN.phase_greg(GP, z[:nb], pl, ocfg, pl_weight=cfg.pl_weight); t['Greg'] = time.perf_counter() - t0
    per_iter = t['Gmain'] + t['Dmain'] + t['Dreg'] / cfg.d_reg_interval + t['Greg'] / cfg.g_reg_interval
    return batch / per_iter, t


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = workload_config(args.workload)
    cores = torch.get_num_threads()
    sample_batch = 2 if cfg.img_resolution <= 256 else 1
    vals = []
    for i in range(args.warmup + args.steps):
        if i < min(args.warmup, 1) or i >= args.warmup:     # one warm-up sample is enough on the CPU
            v, _ = cpu_iteration_rate(cfg, sample_batch)
            if i >= args.warmup:
                vals.append(v)
    value = statistics.mean(vals)
    sample = (f'{args.steps} x (Gmain + Dmain + Dreg + Greg once each at batch {sample_batch}, combined as '
              f'Gmain+Dmain+Dreg/{cfg.d_reg_interval}+Greg/{cfg.g_reg_interval}), fp32')
    line = dict(impl='reference', metric='train img/s (G+D)', value=value, unit='img/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 * sample_batch / value, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD_DESC[args.workload], arm='oracle port of the reference impl=ref path on host CPU'),
                cpu_baseline=dict(value=value, unit='img/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=value, unit='img/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='ffhq256', choices=sorted(WORKLOAD_DESC))
    ap.add_argument('--fp32-mode', default='tf32', choices=['tf32', 'strict'],
                    help="arithmetic of fp32 convolutions: 'tf32' = TF32 tensor cores (torch.backends.cudnn.allow_tf32=True, "
                         "1e-2 class), 'strict' = fp32 FFMA (the reference default perf.allow_tf32=False, 1e-4 class)")
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel eagerly instead of replaying per-phase CUDA graphs')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-roofline', action='store_true', help='skip the eager per-kernel timing loop (ncu launch-list runs)')
    ap.add_argument('--breakdown', default=None, help='write the per-kernel-family time table (json) here')
    args = ap.parse_args()

    if args.impl == 'reference':
        run_reference_arm(args)
        return

    from sgb200 import training, _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    _lib.lib()     # fail loudly if the CUDA library is missing

    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = (args.fp32_mode == 'tf32')
    torch.backends.cuda.matmul.allow_tf32 = (args.fp32_mode == 'tf32')
    cfg = workload_config(args.workload, cuda_graphs=not args.no_graphs)
    tr = training.Trainer(cfg, device, rank=rank, world_size=world)
    R, N = cfg.img_resolution, cfg.batch_gpu
    host_real = torch.randint(0, 256, [N, cfg.img_channels, R, R], dtype=torch.uint8).pin_memory()
    dev_real = host_real.to(device, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed_loop(steps, from_host, profile):
        """returns (ms max over ranks, launches, profile summary or None, last losses)"""
        tr.batch_idx = 0
        barrier()
        if profile:
            _lib.profile_start()
        n0 = _lib.launch_count() + tr.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        nvtx_id = torch.cuda.nvtx.range_start('sgb_timed')   # process-wide range (backward kernels launch from autograd's
                                                              # thread): ncu --nvtx --nvtx-include "sgb_timed" profiles this loop only
        for _ in range(steps):
            real = host_real.to(device, non_blocking=True) if from_host else dev_real
            out = tr.iteration(real, eager=profile)      # per-launch events need eager launches
            if from_host:
                last = {k: float(v) for k, v in out.items()}      # D2H read of every loss of the step
        torch.cuda.synchronize(device)
        torch.cuda.nvtx.range_end(nvtx_id)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() + tr.replayed_launches - n0
        summ = _lib.profile_stop().summary() if profile else None
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, summ, last

    # warm-up: first iteration runs all four phases so that every kernel / allocation is exercised
    for i in range(max(args.warmup, 3)):
        tr.iteration(dev_real, force_all_phases=(i == 0))
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, _, _ = timed_loop(args.steps, from_host=False, profile=False)
    clocks = sampler.stop() if sampler else None
    # per-kernel-family device time (roofline leg): the same K steps launched eagerly with CUDA events around
    # every libsgb200 launch; not part of the reported step time
    summ = None
    if not args.no_roofline:
        _, _, summ, _ = timed_loop(args.steps, from_host=False, profile=True)

    e2e = None
    if not args.no_e2e:
        ms_e, _, _, last = timed_loop(args.steps, from_host=True, profile=False)
        e2e = dict(value=args.steps * N * world / (ms_e / 1000.0), unit='img/s', h2d_bytes_per_step=host_real.numel(),
                   d2h_bytes_per_step=4 * len(last or {}), ms_per_step=ms_e / args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = args.steps * N * world / (ms / 1000.0)
    # dominant kernel family by device time (eager loop, CUDA events around every libsgb200 launch)
    roof = None
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'r1_conv_traffic.json')        # DRAM bytes per launch from ncu (profiles/README.md)
    if os.path.exists(tpath) and args.workload == 'ffhq256':
        with open(tpath) as f:
            traffic = json.load(f)
    if summ:
        total_ms = sum(d['ms'] for d in summ.values()) or 1.0
        dom = max(summ, key=lambda k: summ[k]['ms'])
        d = summ[dom]
        if dom.startswith('conv'):
            # fp32 tensors run TF32 MMAs (half the bf16 rate) in the forward / data-gradient kernels
            tf32 = cfg.num_fp16_res == 0 and args.fp32_mode == 'tf32' and dom.startswith('conv_fwd')
            peak = peaks['tc_sustained'] / (2 if tf32 else 1)
            ach = d['flops'] / (d['ms'] / 1000.0) / 1e12
            roof = dict(kernel=dom, bound='tensor', achieved=ach, peak=peak, unit='TFLOP/s', frac=ach / peak,
                        traffic=(traffic.get(dom) or {}).get('dram_bytes_per_launch'),
                        peak_source=peaks['source'] + (' bf16 sustained / 2 (TF32 MMA rate)' if tf32 else ' bf16 sustained'),
                        launches=d['launches'], avg_launch_ms=d['ms'] / d['launches'], share_of_kernel_time=d['ms'] / total_ms,
                        algorithmic_flops_per_launch=d['flops'] / d['launches'])
        else:
            ach = d['bytes'] / (d['ms'] / 1000.0) / 1e9
            roof = dict(kernel=dom, bound='hbm', achieved=ach, peak=peaks['hbm'], unit='GB/s', frac=ach / peaks['hbm'], traffic=None,
                        peak_source=peaks['source'], launches=d['launches'], avg_launch_ms=d['ms'] / d['launches'],
                        share_of_kernel_time=d['ms'] / total_ms, algorithmic_bytes_per_launch=d['bytes'] / d['launches'])
        if args.breakdown:
            os.makedirs(os.path.dirname(os.path.abspath(args.breakdown)), exist_ok=True)
            with open(args.breakdown, 'w') as f:
                json.dump(dict(step_ms=ms / args.steps, kernel_ms_per_step={k: v['ms'] / args.steps for k, v in summ.items()},
                               detail=summ), f, indent=1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sb = 8 if R <= 256 else 1            # ~10-30 s of CPU work on the box's host cores
        v, t = cpu_iteration_rate(cfg, sb)
        cpu = dict(value=v, unit='img/s', cores=torch.get_num_threads(), kind='port',
                   sample=f'Gmain+Dmain+Dreg/{cfg.d_reg_interval}+Greg/{cfg.g_reg_interval} once each at batch {sb}, fp32 '
                          f'({sum(t.values()):.1f} s of CPU work)')

    line = dict(metric='train img/s (G+D)', value=value, unit='img/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=('tf32' if args.fp32_mode == 'tf32' else 'f32') + ('' if cfg.num_fp16_res == 0 else '+f16'), data='synthetic',
                config=dict(workload=WORKLOAD_DESC[args.workload], batch_per_gpu=N, global_batch=N * world, resolution=R,
                            parallelism=f'dp{world}', g_reg_interval=cfg.g_reg_interval, d_reg_interval=cfg.d_reg_interval,
                            layout='channels_last' if cfg.channels_last else 'nchw', fp32_mode=args.fp32_mode,
                            launch='eager' if args.no_graphs else 'cuda graphs (one per training phase)',
                            l2='working set per step (activations, GBs) far exceeds the 126 MB L2; no explicit flush'),
                gpu_launches=int(launches), e2e=e2e, roofline=roof, cpu_baseline=cpu, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
