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
    """HBM and bf16 tensor peaks: MEASURED_PEAKS.json (driver-written).  The convolutions of the headline workload run on
    kind::tf32 MMAs, for which that file has no number: profiles/r2_tensor_peaks.json holds the TF32 / fp16 rates measured on this
    pool's B200 the same way (benchmarks/measure_peaks.py: cuBLAS 8192^3, burst and sustained)."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        out = dict(hbm=p['hbm_gbs'], tc_burst=p['bf16_tflops'], tc_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                   source='measured (MEASURED_PEAKS.json)')
    else:
        out = dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='fallback (B200_PROFILING.md)')
    tpath = os.path.join(ROOT, 'profiles', 'r2_tensor_peaks.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            t = json.load(f)
        out['tf32_sustained'] = t['tf32']['sustained_tflops']
        out['fp16_sustained'] = t['fp16']['sustained_tflops']
    return out


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
    """bench.py --impl reference: the reference's own CPU implementation of the path (its train_parts modules from the
    snapshot baseline/_ref on CPU tensors => impl='ref' ops + F.conv2d) on the box's host cores, all threads; each step is a
    bounded sample of the workload (every phase once at a reduced batch, combined with the lazy-regularisation weights)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)          # torchrun exports OMP_NUM_THREADS=1: use every host core regardless
    cfg = workload_config(args.workload)
    sample_batch = 2 if cfg.img_resolution <= 256 else 1
    kind = 'reference'
    try:
        from benchmarks import ref_harness
        if ref_harness.reference_root() is None:
            raise RuntimeError('no snapshot')
        fn = lambda: ref_harness.cpu_phase_rate(args.workload, sample_batch)[0]     # noqa: E731
        fn()
        arm = "reference train_parts modules (baseline/_ref snapshot), impl='ref' path on host CPU"
    except Exception:
        kind = 'port'
        fn = lambda: cpu_iteration_rate(cfg, sample_batch)[0]                        # noqa: E731
        fn()
        arm = 'oracle port of the reference impl=ref path on host CPU (reference snapshot not available)'
    vals = [fn() for _ in range(args.steps)]        # the call above was the warm-up sample (one is enough on the CPU)
    value = statistics.mean(vals)
    sample = (f'{args.steps} x (Gmain + Dmain + Dreg + Greg once each at batch {sample_batch}, combined as '
              f'Gmain+Dmain+Dreg/{cfg.d_reg_interval}+Greg/{cfg.g_reg_interval}), fp32, {cores} threads')
    line = dict(impl='reference', metric='train img/s (G+D)', value=value, unit='img/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 * sample_batch / value, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD_DESC[args.workload], arm=arm, sample_batch=sample_batch),
                cpu_baseline=dict(value=value, unit='img/s', cores=cores, kind=kind, sample=sample),
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
    ap.add_argument('--no-secondary', action='store_true', help='skip the config-f 1024^2 run reported as `secondary`')
    ap.add_argument('--no-strict', action='store_true', help='skip the strict-fp32 (reference default arithmetic) run')
    ap.add_argument('--no-callers', action='store_true', help="skip the runs of the reference's unchanged callers (sgb200 / reference GPU path)")
    ap.add_argument('--lean', action='store_true', help='only the headline measurement (ncu launch-list runs)')
    ap.add_argument('--start-idx', type=int, default=0, help='iteration index the timed loop starts at (ncu launch lists: 1 = an '
                    'iteration with the main phases only; the default 0 = whole lazy-regularisation periods)')
    args = ap.parse_args()

    if args.lean:
        args.no_secondary = args.no_strict = args.no_callers = args.no_cpu_baseline = args.no_e2e = args.no_roofline = True
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    from sgb200 import _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        # the in-graph gradient all-reduce (~120 MB per phase): more NCCL channels than its default for a message of this size
        # (measured on 2 B200s, benchmarks/gpu_runs/r2_run38.sh: 69.77 -> 69.38 ms per step); an explicit setting wins
        os.environ.setdefault('NCCL_MIN_NCHANNELS', '32')
        dist.init_process_group('nccl', device_id=device)
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    _lib.lib()     # fail loudly if the CUDA library is missing
    torch.backends.cudnn.benchmark = False
    ctx = dict(world=world, rank=rank, local_rank=local_rank, device=device, peaks=load_peaks())

    # 1) the headline: BASELINE.json configs[1] (or --workload), fp32 tensors on the arithmetic --fp32-mode names
    main_m = measure(ctx, args, args.workload, args.fp32_mode, args.steps, e2e=not args.no_e2e, roofline=not args.no_roofline,
                     clocks=True, breakdown=args.breakdown)
    # 2) config C on the same record: config-f 1024^2, 4 img/GPU, fp16 top-4 resolutions, one lazy-regularisation period
    secondary = None
    if not args.no_secondary and args.workload != 'f1024':
        m = measure(ctx, args, 'f1024', args.fp32_mode, 16, e2e=not args.no_e2e, roofline=not args.no_roofline, clocks=False,
                    breakdown=(args.breakdown + '.f1024.json') if args.breakdown else None)
        if rank == 0:
            secondary = dict(workload=WORKLOAD_DESC['f1024'], metric='train img/s (G+D)', value=m['value'], unit='img/s',
                             ms_per_step=m['ms_per_step'], steps=16, batch_per_gpu=m['batch'], n_gpus=world, dtype=m['dtype'],
                             gpu_launches=m['launches'], e2e=m['e2e'], roofline=m['roofline'], families=m['families'])
    # 3) the reference-default arithmetic (perf.allow_tf32 = False): strict fp32 on the FFMA kernels, same workload
    strict = None
    if not args.no_strict and args.fp32_mode != 'strict' and world == 1:
        m = measure(ctx, args, args.workload, 'strict', min(args.steps, 16), e2e=False, roofline=False, clocks=False)
        strict = dict(value=m['value'], unit='img/s', ms_per_step=m['ms_per_step'], steps=min(args.steps, 16), dtype='f32',
                      note='torch.backends.cudnn.allow_tf32=False (reference default, arguments.py:78): fp32 FFMA convolution kernels')
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # 4) the reference's UNCHANGED callers (train_parts G / D / SG2Loss / R1reg / PPLreg from baseline/_ref): on the sgb200
    #    kernels through sgb200.install(), and on the reference's own CUDA path (its JIT plugins + cuDNN) = the kernel to beat
    callers = vs_ref_gpu = None
    if not args.no_callers and world == 1:
        callers = run_ref_harness('sgb200', args.workload, args.fp32_mode)
        ref_gpu = run_ref_harness('reference', args.workload, args.fp32_mode)
        if isinstance(ref_gpu, dict) and 'img_per_s' in ref_gpu:
            vs_ref_gpu = dict(reference_gpu_img_per_s=ref_gpu['img_per_s'], reference_gpu_ms_per_step=ref_gpu['ms_per_step'],
                              ratio_value=main_m['value'] / ref_gpu['img_per_s'],
                              ratio_unchanged_callers=(callers['img_per_s'] / ref_gpu['img_per_s']) if isinstance(callers, dict) and 'img_per_s' in callers else None,
                              what="reference's own CUDA path on this GPU: bias_act_plugin / upfirdn2d_plugin (custom_ops.py JIT, sm_100a) + "
                                   "cuDNN via F.conv2d, eager launches, cudnn.benchmark=True, same workload / batch / fp32 mode")
        else:
            vs_ref_gpu = ref_gpu
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(args.workload)

    cfg = main_m['cfg']
    line = dict(metric='train img/s (G+D)', value=main_m['value'], unit='img/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=main_m['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype=main_m['dtype'],
                data='synthetic',
                config=dict(workload=WORKLOAD_DESC[args.workload], batch_per_gpu=main_m['batch'], global_batch=main_m['batch'] * world,
                            resolution=cfg.img_resolution, parallelism=f'dp{world}', g_reg_interval=cfg.g_reg_interval,
                            d_reg_interval=cfg.d_reg_interval, layout='channels_last' if cfg.channels_last else 'nchw',
                            fp32_mode=args.fp32_mode, launch='eager' if args.no_graphs else 'cuda graphs (one per training phase)',
                            l2='working set per step (activations, GBs) far exceeds the 126 MB L2; no explicit flush'),
                gpu_launches=int(main_m['launches']), e2e=main_m['e2e'], roofline=main_m['roofline'], families=main_m['families'],
                cpu_baseline=cpu, clocks=main_m['clocks'], secondary=secondary, strict_fp32=strict, callers_reference=callers,
                vs_reference_gpu=vs_ref_gpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ref_harness(backend, workload, fp32_mode, steps=16, timeout=420):
    """benchmarks/ref_harness.py in its own process (the two backends bind the same module names) -> its JSON line."""
    cmd = [sys.executable, os.path.join(ROOT, 'benchmarks', 'ref_harness.py'), '--backend', backend, '--workload', workload,
           '--fp32-mode', fp32_mode, '--steps', str(steps), '--warmup', '3']
    if backend == 'reference':
        cmd.append('--cudnn-benchmark')
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        return dict(unavailable=(r.stderr.strip().splitlines() or ['no output'])[-1][:300])
    except Exception as e:     # missing snapshot, plugin build failure, time-out: report, never fail the bench
        return dict(unavailable=f'{type(e).__name__}: {e}'[:300])


def measure(ctx, args, workload, fp32_mode, steps, e2e, roofline, clocks, breakdown=None):
    """One workload on this rank's GPU: W warm-up iterations, exactly `steps` timed iterations (barrier + synchronize on both
    sides, max over ranks), optionally the end-to-end loop and the per-kernel-family eager loop."""
    from sgb200 import training, _lib
    world, rank, device, peaks = ctx['world'], ctx['rank'], ctx['device'], ctx['peaks']
    torch.backends.cudnn.allow_tf32 = (fp32_mode == 'tf32')
    torch.backends.cuda.matmul.allow_tf32 = (fp32_mode == 'tf32')
    cfg = workload_config(workload, cuda_graphs=not args.no_graphs, allreduce_in_graph=os.environ.get('SGB_ALLREDUCE_IN_GRAPH', '1') != '0')
    tr = training.Trainer(cfg, device, rank=rank, world_size=world)
    R, N = cfg.img_resolution, cfg.batch_gpu
    host_real = torch.randint(0, 256, [N, cfg.img_channels, R, R], dtype=torch.uint8).pin_memory()
    dev_real = host_real.to(device, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed_loop(nsteps, from_host, profile):
        """returns (ms max over ranks, launches, profile summary or None, last losses)"""
        tr.batch_idx = args.start_idx
        barrier()
        if profile:
            _lib.profile_start()
        n0 = _lib.launch_count() + tr.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        nvtx_id = torch.cuda.nvtx.range_start('sgb_timed')   # process-wide range (backward kernels launch from autograd's
                                                              # thread): ncu --nvtx --nvtx-include "sgb_timed" profiles this loop only
        for _ in range(nsteps):
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
        prof = _lib.profile_stop() if profile else None
        summ = (prof.summary(), prof.summary(by_tag=True)) if profile else None
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, summ, last

    # warm-up: first iteration runs all four phases so that every kernel / allocation is exercised
    for i in range(max(args.warmup, 3)):
        tr.iteration(dev_real, force_all_phases=(i == 0))
    barrier()
    sampler = ClockSampler(ctx['local_rank']) if (rank == 0 and clocks) else None
    if sampler:
        sampler.start()
    ms, launches, _, _ = timed_loop(steps, from_host=False, profile=False)
    clk = sampler.stop() if sampler else None
    # per-kernel-family device time (roofline leg): the same K steps launched eagerly with CUDA events around
    # every libsgb200 launch; not part of the reported step time
    summ = None
    if roofline:
        _, _, summ, _ = timed_loop(steps, from_host=False, profile=True)
        summ, by_tag = summ
    e2e_d = None
    if e2e:
        ms_e, _, _, last = timed_loop(steps, from_host=True, profile=False)
        e2e_d = dict(value=steps * N * world / (ms_e / 1000.0), unit='img/s', h2d_bytes_per_step=host_real.numel(),
                     d2h_bytes_per_step=4 * len(last or {}), ms_per_step=ms_e / steps)
    out = dict(cfg=cfg, batch=N, value=steps * N * world / (ms / 1000.0), ms_per_step=ms / steps, launches=int(launches), e2e=e2e_d,
               clocks=clk, roofline=None, families=None,
               dtype=('tf32' if fp32_mode == 'tf32' else 'f32') + ('' if cfg.num_fp16_res == 0 else '+f16'))
    if rank == 0 and summ:
        out['roofline'], out['families'] = roofline_of(summ, cfg, workload, fp32_mode, peaks, steps)
        if breakdown:
            os.makedirs(os.path.dirname(os.path.abspath(breakdown)), exist_ok=True)
            with open(breakdown, 'w') as f:
                json.dump(dict(step_ms=ms / steps, kernel_ms_per_step={k: v['ms'] / steps for k, v in summ.items()}, detail=summ,
                               by_shape=by_tag), f, indent=1)
    del tr
    torch.cuda.empty_cache()
    return out


def roofline_of(summ, cfg, workload, fp32_mode, peaks, steps):
    """(roofline object of the dominant kernel family, {family: ms per step / achieved / fraction}) from the CUDA-event summary.
    Tensor peaks: MEASURED_PEAKS.json has the bf16 rate only; fp16 MMAs run at the bf16 rate and kind::tf32 MMAs at half of it
    (tcgen05 K = 8 per instruction instead of 16), so fp32 tensors are reported against bf16 sustained / 2.  Mixed workloads
    (f1024: fp32 below 128^2, fp16 above) are reported against the fp16 rate, the stricter denominator."""
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'r2_conv_traffic.json')        # DRAM bytes per launch from ncu (profiles/README.md)
    if os.path.exists(tpath) and workload == 'ffhq256':
        with open(tpath) as f:
            traffic = json.load(f)
    total_ms = sum(d['ms'] for d in summ.values()) or 1.0
    tf32 = cfg.num_fp16_res == 0 and fp32_mode == 'tf32'
    if tf32 and 'tf32_sustained' in peaks:
        tc_peak, tc_src = peaks['tf32_sustained'], 'measured cuBLAS TF32 8192^3 sustained (profiles/r2_tensor_peaks.json, benchmarks/measure_peaks.py)'
    elif tf32:
        tc_peak, tc_src = peaks['tc_sustained'] / 2, peaks['source'] + ' bf16 sustained / 2 (TF32 MMA rate)'
    else:
        tc_peak, tc_src = peaks['tc_sustained'], peaks['source'] + ' bf16 sustained (= fp16 MMA rate)'
    fam = {}
    for k, d in summ.items():
        e = dict(ms_per_step=d['ms'] / steps, launches_per_step=d['launches'] / steps)
        if k.startswith('conv') and d['flops'] > 0:
            e['tflops'] = d['flops'] / (d['ms'] / 1000.0) / 1e12
            e['frac_tensor'] = e['tflops'] / tc_peak
        if d['bytes'] > 0:
            e['gbs'] = d['bytes'] / (d['ms'] / 1000.0) / 1e9
            e['frac_hbm'] = e['gbs'] / peaks['hbm']
        fam[k] = e
    dom = max(summ, key=lambda k: summ[k]['ms'])
    d = summ[dom]
    if dom.startswith('conv'):
        ach = d['flops'] / (d['ms'] / 1000.0) / 1e12
        roof = dict(kernel=dom, bound='tensor', achieved=ach, peak=tc_peak, unit='TFLOP/s', frac=ach / tc_peak,
                    traffic=(traffic.get(dom) or {}).get('dram_bytes_per_launch'),
                    peak_source=tc_src,
                    launches=d['launches'], avg_launch_ms=d['ms'] / d['launches'], share_of_kernel_time=d['ms'] / total_ms,
                    algorithmic_flops_per_launch=d['flops'] / d['launches'],
                    hbm_view=dict(gbs=d['bytes'] / (d['ms'] / 1000.0) / 1e9, frac=d['bytes'] / (d['ms'] / 1000.0) / 1e9 / peaks['hbm']))
    else:
        ach = d['bytes'] / (d['ms'] / 1000.0) / 1e9
        roof = dict(kernel=dom, bound='hbm', achieved=ach, peak=peaks['hbm'], unit='GB/s', frac=ach / peaks['hbm'], traffic=None,
                    peak_source=peaks['source'], launches=d['launches'], avg_launch_ms=d['ms'] / d['launches'],
                    share_of_kernel_time=d['ms'] / total_ms, algorithmic_bytes_per_launch=d['bytes'] / d['launches'])
    return roof, fam


def cpu_baseline(workload):
    """The reference's own modules (baseline/_ref, impl='ref' on CPU tensors) on this box's host cores, all threads, on a
    bounded sample: every phase once at a reduced batch, combined with the lazy-regularisation weights.  Falls back to the
    oracle port (pinned to the reference by tests/golden) when the snapshot is absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    R = {'ffhq256': 256, 'f1024': 1024, 'sg2ada64': 64}[workload]
    sb = 8 if R <= 256 else 1
    try:
        from benchmarks import ref_harness
        if ref_harness.reference_root() is None:
            raise RuntimeError('no snapshot')
        v, t, gi, di = ref_harness.cpu_phase_rate(workload, sb)
        return dict(value=v, unit='img/s', cores=cores, kind='reference',
                    sample=f"reference train_parts G/D + SG2Loss/R1reg/PPLreg (baseline/_ref, impl='ref'), Gmain+Dmain+Dreg/{di}+Greg/{gi} "
                           f'once each at batch {sb}, fp32 ({sum(t.values()):.1f} s of CPU work)')
    except Exception as e:
        cfg = workload_config(workload)
        v, t = cpu_iteration_rate(cfg, sb)
        return dict(value=v, unit='img/s', cores=cores, kind='port',
                    sample=f'oracle port ({type(e).__name__}: reference snapshot unusable), Gmain+Dmain+Dreg/{cfg.d_reg_interval}+'
                           f'Greg/{cfg.g_reg_interval} once each at batch {sb}, fp32 ({sum(t.values()):.1f} s of CPU work)')


if __name__ == '__main__':
    main()
