"""Drive the reference's UNCHANGED callers (train_parts/generators.py, discriminators.py, losses_base.py,
regularizations.py, optimizers.py) from the snapshot under `baseline/_ref/` (baseline/snapshot_reference.py), either

  backend='sgb200'     after `sgb200.install()`: the reference's G / D / SG2Loss / R1reg / PPLreg run on the libsgb200
                       kernels (what a user who switches over gets without touching the reference), or
  backend='reference'  on the reference's own path: CUDA tensors -> its JIT-built `bias_act_plugin` /
                       `upfirdn2d_plugin` + cuDNN through F.conv2d ("the kernel to beat"); CPU tensors -> impl='ref'
                       (the CPU baseline of bench.py --impl reference).

One backend per process (both rebind the same module names).  Test / benchmark infrastructure: nothing under
`style-big-gan_b200/` imports this file, and nothing here reads `/root/reference` (it does not exist on the GPU box).

The training iteration below restates the reference's loop body (train_parts/trainers.py:699-761) around the
reference's own `loss.accumulate_gradients`: phases Gmain / Greg / Dmain / Dreg with lazy regularisation
(:601-626), `misc.nan_to_num` on every gradient + Adam step (:745-748), G_ema lerp (:752-761).
"""
import copy
import dataclasses
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.realpath('/root/repo') == os.path.realpath(ROOT):
    ROOT = '/root/repo'       # same path in the authoring container and on the GPU box (a symlink there): the plugin build
                              # directory prepared by build() is then found again, not rebuilt
_backend = None


def reference_root():
    """The travelling snapshot (or $SGB_REFERENCE_ROOT); None when absent."""
    for p in (os.environ.get('SGB_REFERENCE_ROOT'), os.path.join(ROOT, 'baseline', '_ref')):
        if p and os.path.isfile(os.path.join(p, 'train_parts', 'generators.py')):
            return p
    return None


def _shims():
    """the reference imports omegaconf (absent here) and relies on pre-3.10 stdlib behaviour."""
    if 'omegaconf' not in sys.modules:
        om = types.ModuleType('omegaconf')
        om.MISSING = '???'
        om.OmegaConf = type('OmegaConf', (), {})
        lc = types.ModuleType('omegaconf.listconfig')
        lc.ListConfig = list
        om.listconfig = lc
        sys.modules['omegaconf'] = om
        sys.modules['omegaconf.listconfig'] = lc
    if not getattr(dataclasses.make_dataclass, '_sgb_shim', False):
        orig = dataclasses.make_dataclass

        def make_dataclass(*a, **k):
            k.setdefault('eq', False)
            return orig(*a, **k)
        make_dataclass._sgb_shim = True
        dataclasses.make_dataclass = make_dataclass
    import collections
    import collections.abc
    if not hasattr(collections, 'MutableMapping'):
        collections.MutableMapping = collections.abc.MutableMapping


def _plugin_import_shim():
    """The reference's custom_ops.get_plugin (custom_ops.py:109-111) builds a plugin with torch.utils.cpp_extension.load and then
    calls importlib.import_module(name).  torch 1.7 registered the built module in sys.modules; torch 2.x does not, so on this
    image the import raises, the op prints "Failed!" and silently falls back to its slow impl='ref' path -- which would make
    the reference's GPU numbers meaningless as "the kernel to beat".  Register the module the way torch 1.7 did."""
    import torch.utils.cpp_extension as ce
    if getattr(ce.load, '_sgb_shim', False):
        return
    orig = ce.load

    def load(name, *a, **k):
        mod = orig(name, *a, **k)
        if isinstance(mod, types.ModuleType):
            sys.modules.setdefault(name, mod)
        return mod
    load._sgb_shim = True
    ce.load = load


def plugin_status():
    """{'bias_act': True/False, 'upfirdn2d': True/False}: did the reference's CUDA plugins load (backend='reference')?"""
    from stylegan2ada.torch_utils.ops import bias_act as rb, upfirdn2d as ru
    return dict(bias_act=bool(rb._init()), upfirdn2d=bool(ru._init()))


def import_reference(backend, repair_plugins=True):
    """Put the snapshot on sys.path (and sgb200 in front of it for backend='sgb200').  Returns the root."""
    global _backend
    assert backend in ('sgb200', 'reference')
    root = reference_root()
    if root is None:
        raise RuntimeError('no reference snapshot: run `python baseline/snapshot_reference.py` where /root/reference exists '
                           '(build() does it) -- baseline/_ref/ travels to the GPU box with gpurun')
    if _backend is not None and _backend != backend:
        raise RuntimeError(f'this process already runs the reference callers on backend {_backend!r}')
    _shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    if backend == 'sgb200':
        pkg = os.path.join(ROOT, 'style-big-gan_b200')
        if pkg not in sys.path:
            sys.path.insert(0, pkg)
        import sgb200
        sgb200.install()
    else:
        # the plugins are JIT-built by the reference's own custom_ops.get_plugin; keep the build inside the snapshot so a
        # build made in one gpurun call is found again and this container's ~/.cache is not involved
        os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(root, '_torch_extensions'))
        os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
        if repair_plugins:
            _plugin_import_shim()
    _backend = backend
    return root


# ---------------------------------------------------------------------------------------------------------------
# The four BASELINE.json training configurations as the reference's own keyword arguments
# (configs/*.yaml -> arguments.py -> trainers.py:537-541; config-f = stylegan2ada/train.py:156,181-182).
WORKLOADS = {
    'sg2ada64': dict(res=64, batch_gpu=8, z_dim=512, w_dim=512, map_layers=2, channel_base=32768, d_arch='orig', mbstd=32,
                     r1_gamma=0.01, ppl=False, style_mixing_prob=0.0, num_fp16_res=0, conv_clamp=None, ema_kimg=500.0,
                     g_attn=(), d_attn=()),
    'ffhq256': dict(res=256, batch_gpu=32, z_dim=512, w_dim=512, map_layers=6, channel_base=16384, d_arch='resnet', mbstd=8,
                    r1_gamma=1.0, ppl=True, style_mixing_prob=0.9, num_fp16_res=0, conv_clamp=None, ema_kimg=20.0,
                    g_attn=(), d_attn=()),
    'f1024': dict(res=1024, batch_gpu=4, z_dim=512, w_dim=512, map_layers=8, channel_base=32768, d_arch='resnet', mbstd=4,
                  r1_gamma=10.0, ppl=True, style_mixing_prob=0.9, num_fp16_res=4, conv_clamp=256, ema_kimg=10.0,
                  g_attn=(), d_attn=()),
    'sg2attent256': dict(res=256, batch_gpu=64, z_dim=512, w_dim=512, map_layers=2, channel_base=32768, d_arch='orig', mbstd=32,
                         r1_gamma=0.01, ppl=False, style_mixing_prob=0.0, num_fp16_res=4, conv_clamp=256, ema_kimg=500.0,
                         g_attn=(32, 16, 8, 4), d_attn=(32,)),
}


def _to_easy(obj, dnnlib):
    if dataclasses.is_dataclass(obj):
        return dnnlib.EasyDict({f.name: _to_easy(getattr(obj, f.name), dnnlib) for f in dataclasses.fields(obj)})
    if isinstance(obj, dict):
        return dnnlib.EasyDict({k: _to_easy(v, dnnlib) for k, v in obj.items()})
    return obj


def build_nets(w, device, channel_max=512):
    """The reference's registered 'sg2_classic' G and D (train_parts/generators.py:533, discriminators.py:402)."""
    import stylegan2ada.dnnlib as dnnlib
    from train_parts.generators import generators
    from train_parts.discriminators import discriminators
    gk = _to_easy(generators.args['sg2_classic'](), dnnlib)
    gk.update(z_dim=w['z_dim'], w_dim=w['w_dim'], c_dim=0, img_resolution=w['res'], img_channels=3, attentions=tuple(w['g_attn']))
    gk.mapping_kwargs.num_layers = w['map_layers']
    gk.synthesis_kwargs.channel_base = w['channel_base']
    gk.synthesis_kwargs.channel_max = channel_max
    gk.synthesis_kwargs.num_fp16_res = w['num_fp16_res']
    gk.synthesis_kwargs.block_kwargs.conv_clamp = w['conv_clamp']
    G = generators['sg2_classic'](**gk)
    dk = _to_easy(discriminators.args['sg2_classic'](), dnnlib)
    dk.update(c_dim=0, img_resolution=w['res'], img_channels=3, architecture=w['d_arch'], channel_base=w['channel_base'],
              channel_max=channel_max, num_fp16_res=w['num_fp16_res'], conv_clamp=w['conv_clamp'], attentions=tuple(w['d_attn']))
    dk.epilogue_kwargs.mbstd_group_size = w['mbstd']
    D = discriminators['sg2_classic'](**dk)
    return G.to(device), D.to(device)


class _Synthesis(torch.nn.Module):
    """`G_synthesis(ws)` of SG2Loss.run_G with a fixed noise_mode (the reference passes no kwargs => 'random')."""

    def __init__(self, syn, noise_mode):
        super().__init__()
        self.syn, self.noise_mode = syn, noise_mode

    def forward(self, ws):
        return self.syn(ws, noise_mode=self.noise_mode)


class RefCallerTrainer:
    """One rank's training state built from the reference's own classes; `iteration(real_u8)` = one loop body of
    BaseTrainer.training_loop (trainers.py:710-761) at batch == batch_gpu (one accumulation round)."""

    def __init__(self, workload, device, backend, seed=0, noise_mode='random', g_reg_interval=16, d_reg_interval=4, lr=0.0025,
                 betas=(0.0, 0.99), channel_max=512, use_ema=True, **over):
        import_reference(backend)
        import stylegan2ada.dnnlib as dnnlib
        from stylegan2ada.torch_utils import misc
        from train_parts.losses_base import losses_arch
        from train_parts.optimizers import optimizers
        self.w = dict(WORKLOADS[workload] if isinstance(workload, str) else workload)
        self.w.update(over)
        w = self.w
        self.device, self.misc = torch.device(device), misc
        torch.manual_seed(seed)
        self.G, self.D = build_nets(w, self.device, channel_max)
        self.G.train().requires_grad_(False)
        self.D.train().requires_grad_(False)
        self.G_ema = copy.deepcopy(self.G).eval() if use_ema else None
        syn = self.G.synthesis if noise_mode == 'random' else _Synthesis(self.G.synthesis, noise_mode)
        gen_regs = [('ppl', dict(pl_batch_shrink=2, pl_decay=0.01, pl_weight=2.0))] if w['ppl'] else []
        dis_regs = [('r1', dict(r1_gamma=w['r1_gamma']))] if w['r1_gamma'] else []
        self.loss = losses_arch['sg2'](G_mapping=self.G.mapping, G_synthesis=syn, D=self.D, device=self.device, gen_regs=gen_regs,
                                       dis_regs=dis_regs, loss='softplus', style_mixing_prob=w['style_mixing_prob'])
        self.phases = []
        for name, module, interval, has_reg in [('G', self.G, g_reg_interval, bool(gen_regs)), ('D', self.D, d_reg_interval, bool(dis_regs))]:
            if not has_reg:         # the reference would build a 'Gboth' phase for reg_interval <= 0; with no regulariser
                opt = optimizers['adam'](params=module.parameters(), lr=lr, betas=list(betas), eps=1e-8)      # it equals 'Gmain'
                self.phases.append(dnnlib.EasyDict(name=name + 'main', module=module, opt=opt, interval=1))
            else:                   # lazy regularisation, trainers.py:612-620
                r = interval / (interval + 1)
                opt = optimizers['adam'](params=module.parameters(), lr=lr * r, betas=[b ** r for b in betas], eps=1e-8)
                self.phases.append(dnnlib.EasyDict(name=name + 'main', module=module, opt=opt, interval=1))
                self.phases.append(dnnlib.EasyDict(name=name + 'reg', module=module, opt=opt, interval=interval))
        self.batch_idx = 0
        self.c = torch.zeros([w['batch_gpu'], 0], device=self.device)

    def phase_grads(self, name, real, z, gain):
        """zero_grad + requires_grad toggling + the reference's accumulate_gradients (trainers.py:733-742)."""
        ph = next(p for p in self.phases if p.name == name)
        ph.opt.zero_grad(set_to_none=True)
        ph.module.requires_grad_(True)
        self.loss.accumulate_gradients(phase=name, real_img=real, real_c=self.c[:real.shape[0]], gen_z=z, gen_c=self.c[:z.shape[0]],
                                       sync=True, gain=gain)
        ph.module.requires_grad_(False)
        return ph

    def iteration(self, real_u8, force_all_phases=False):
        w = self.w
        real = real_u8.to(self.device).to(torch.float32) / 127.5 - 1
        zs = torch.randn([len(self.phases), w['batch_gpu'], w['z_dim']], device=self.device)
        ran = []
        for ph, z in zip(self.phases, zs):
            if not force_all_phases and self.batch_idx % ph.interval != 0:
                continue
            self.phase_grads(ph.name, real, z, ph.interval)
            for p in ph.module.parameters():
                if p.grad is not None:
                    self.misc.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
            ph.opt.step()
            ran.append(ph.name)
        if self.G_ema is not None:
            beta = 0.5 ** (w['batch_gpu'] / max(w['ema_kimg'] * 1000, 1e-8))
            with torch.no_grad():
                for p_ema, p in zip(self.G_ema.parameters(), self.G.parameters()):
                    p_ema.copy_(p.lerp(p_ema, beta))
                for b_ema, b in zip(self.G_ema.buffers(), self.G.buffers()):
                    b_ema.copy_(b)
        self.batch_idx += 1
        return ran


def time_iterations(tr, steps, warmup, device):
    """ms per iteration of `tr.iteration` (CUDA events; wall clock on CPU); K steps from batch_idx 0 = whole lazy-reg periods."""
    import time
    w = tr.w
    real = torch.randint(0, 256, [w['batch_gpu'], 3, w['res'], w['res']], dtype=torch.uint8, device=device)
    for i in range(warmup):
        tr.iteration(real, force_all_phases=(i == 0))
    tr.batch_idx = 0
    if torch.device(device).type == 'cuda':
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.iteration(real)
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(real)
    return (time.perf_counter() - t0) * 1000.0 / steps


def cpu_phase_rate(workload, batch, g_reg_interval=16, d_reg_interval=4):
    """CPU baseline sample: the reference's own phases (impl='ref' ops, F.conv2d) once each at per-step batch `batch`
    (Greg on batch // 2, regularizations.py:22), combined into the amortised iteration Gmain + Dmain + Dreg/di + Greg/gi.
    Returns (img/s, {phase: seconds}, gi, di)."""
    import time
    tr = RefCallerTrainer(workload, 'cpu', 'reference', use_ema=False, g_reg_interval=g_reg_interval, d_reg_interval=d_reg_interval,
                          batch_gpu=batch, mbstd=min(WORKLOADS[workload]['mbstd'], batch))
    w = tr.w
    g = torch.Generator().manual_seed(2)
    z = torch.randn(batch, w['z_dim'], generator=g)
    real = torch.rand(batch, 3, w['res'], w['res'], generator=g) * 2 - 1
    t = dict(Gmain=0.0, Dmain=0.0, Greg=0.0, Dreg=0.0)
    for ph in tr.phases:
        t0 = time.perf_counter()
        tr.phase_grads(ph.name, real, z, ph.interval)
        t[ph.name] = time.perf_counter() - t0
    per_iter = t['Gmain'] + t['Dmain'] + t['Dreg'] / d_reg_interval + t['Greg'] / g_reg_interval
    return batch / per_iter, t, g_reg_interval, d_reg_interval


def main():
    """python benchmarks/ref_harness.py --backend sgb200|reference --workload ffhq256 [--tf32] -> one JSON line"""
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument('--backend', required=True, choices=['sgb200', 'reference'])
    ap.add_argument('--workload', default='ffhq256', choices=sorted(WORKLOADS))
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--fp32-mode', default='tf32', choices=['tf32', 'strict'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--cudnn-benchmark', action='store_true')
    ap.add_argument('--as-is', action='store_true', help="backend=reference without the plugin import repair: what the unmodified "
                    "reference does on torch 2.x (plugins fail to import, ops fall back to impl='ref')")
    a = ap.parse_args()
    if a.backend == 'reference':
        import_reference('reference', repair_plugins=not a.as_is)
    tf32 = a.fp32_mode == 'tf32'
    torch.backends.cudnn.allow_tf32 = tf32             # trainers.py:510-511 (perf.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = a.cudnn_benchmark  # trainers.py:509 (perf.cudnn_benchmark)
    over = {} if a.batch is None else dict(batch_gpu=a.batch)
    if a.device == 'cpu':
        torch.set_num_threads(os.cpu_count() or 1)
    tr = RefCallerTrainer(a.workload, a.device, a.backend, **over)
    ms = time_iterations(tr, a.steps, a.warmup, a.device)
    n = tr.w['batch_gpu']
    plugins = plugin_status() if (a.backend == 'reference' and a.device != 'cpu') else None
    print(json.dumps(dict(callers='reference (unchanged train_parts)', backend=a.backend, device=a.device, workload=a.workload,
                          batch_gpu=n, fp32_mode=a.fp32_mode, steps=a.steps, ms_per_step=ms, img_per_s=1000.0 * n / ms,
                          cudnn_benchmark=a.cudnn_benchmark, threads=torch.get_num_threads(), reference_cuda_plugins=plugins)), flush=True)


if __name__ == '__main__':
    main()
