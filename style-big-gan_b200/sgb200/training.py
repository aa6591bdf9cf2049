"""Training-step harness for the StyleGAN2 G / D pair on the sgb200 ops (synthetic data; used by bench.py,
smoke() and the tests).  It reproduces what the reference trainer does per iteration, nothing else:

  phases and lazy regularisation      train_parts/trainers.py:601-633  (Gmain, Greg every g_reg_interval,
                                       Dmain, Dreg every d_reg_interval; lr and betas rescaled by
                                       interval / (interval + 1))
  per-phase work                      train_parts/losses_base.py:43-109 (SG2Loss.run_G with style mixing :131-141)
  path-length regulariser             train_parts/regularizations.py:11-37
  R1 regulariser                      train_parts/regularizations.py:40-56
  non-saturating logistic loss        train_parts/losses.py:47-58 ('softplus')
  update                              train_parts/trainers.py:745-748 (nan_to_num on grads, Adam step)
  G_ema                               train_parts/trainers.py:752-761

Data parallelism: one process per GPU; every rank holds full replicas and processes `batch_gpu` images per
phase; parameter gradients are averaged with ONE flat all-reduce per phase (NCCL over NVLink), which is what
the reference's DistributedDataParallel wrapping amounts to (trainers.py:883-893) -- the hot path itself has no
collective.  Logging, snapshots, metrics, augmentation and datasets are out of scope.
"""
import copy
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from . import networks
from .ops import conv2d_gradfix
from .ops import fused_conv


@dataclass
class TrainConfig:
    img_resolution: int = 256
    img_channels: int = 3
    z_dim: int = 512
    w_dim: int = 512
    channel_base: int = 16384
    channel_max: int = 512
    map_layers: int = 6
    num_fp16_res: int = 0
    conv_clamp: float = None
    d_arch: str = 'resnet'
    mbstd_group_size: int = 8
    batch_gpu: int = 32
    lr: float = 0.0025
    betas: tuple = (0.0, 0.99)
    g_reg_interval: int = 16
    d_reg_interval: int = 4
    r1_gamma: float = 1.0
    pl_weight: float = 2.0
    pl_batch_shrink: int = 2
    pl_decay: float = 0.01
    style_mixing_prob: float = 0.9
    use_ppl: bool = True
    use_r1: bool = True
    ema_kimg: float = 20.0
    use_ema: bool = True
    channels_last: bool = True
    noise_mode: str = 'random'
    seed: int = 0
    force_flat_grads: bool = False    # tests: use the flat gradient buffers (the N > 1 path) on a single GPU
    allreduce_in_graph: bool = True   # N > 1 with cuda_graphs: capture the NCCL gradient all-reduce inside each phase graph
    cuda_graphs: bool = False      # capture each training phase in a CUDA graph and replay it (static shapes; removes the
                                   # per-launch host cost of ~3000 kernel launches per iteration)


# named workloads of BASELINE.json `configs`
def config_ffhq256(**over):
    """configs/ffhq_sg2.yaml: 256x256, batch 32/GPU, channel_base 16384, 6 mapping layers, R1 gamma 1 + PPL."""
    return TrainConfig(**over)


def config_f1024(**over):
    """stylegan2ada/train.py:156,181-182 'stylegan2' preset (config-f): 1024x1024, 4/GPU, fp16 top-4, clamp 256."""
    kw = dict(img_resolution=1024, channel_base=32768, map_layers=8, num_fp16_res=4, conv_clamp=256.0, batch_gpu=4,
              mbstd_group_size=4, r1_gamma=10.0, ema_kimg=10.0)
    kw.update(over)
    return TrainConfig(**kw)


def config_sg2ada64(**over):
    """configs/sg2ada.yaml at 64x64 batch 8: 2 mapping layers, D 'orig', mbstd 32, R1 gamma 0.01, no PPL, no style mixing."""
    kw = dict(img_resolution=64, channel_base=32768, map_layers=2, d_arch='orig', mbstd_group_size=32, batch_gpu=8,
              r1_gamma=0.01, use_ppl=False, style_mixing_prob=0.0, ema_kimg=500.0)
    kw.update(over)
    return TrainConfig(**kw)


def build_networks(cfg, device):
    G = networks.Generator(
        z_dim=cfg.z_dim, c_dim=0, w_dim=cfg.w_dim, img_resolution=cfg.img_resolution, img_channels=cfg.img_channels,
        mapping_kwargs=dict(num_layers=cfg.map_layers),
        synthesis_kwargs=dict(channel_base=cfg.channel_base, channel_max=cfg.channel_max, num_fp16_res=cfg.num_fp16_res,
                              conv_clamp=cfg.conv_clamp, channels_last=cfg.channels_last))
    D = networks.Discriminator(
        c_dim=0, img_resolution=cfg.img_resolution, img_channels=cfg.img_channels, architecture=cfg.d_arch,
        channel_base=cfg.channel_base, channel_max=cfg.channel_max, num_fp16_res=cfg.num_fp16_res, conv_clamp=cfg.conv_clamp,
        block_kwargs=dict(channels_last=cfg.channels_last), epilogue_kwargs=dict(mbstd_group_size=cfg.mbstd_group_size))
    return G.to(device), D.to(device)


def nan_to_num_(tensors, nan=0.0, posinf=1e5, neginf=-1e5):
    """In-place `torch.nan_to_num` of every tensor of the list (train_parts/trainers.py:745-748 does it per parameter
    gradient): dense fp32 CUDA tensors share 1-2 launches (`sgb_nan_to_num_multi`), anything else goes one by one."""
    import ctypes
    from . import _lib
    fast = [t for t in tensors if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()]
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            torch.nan_to_num(t, nan=nan, posinf=posinf, neginf=neginf, out=t)
    by_dev = {}
    for t in fast:
        by_dev.setdefault(t.device, []).append(t)
    for dev, ts in by_dev.items():
        ptrs = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        nums = (ctypes.c_int64 * len(ts))(*[t.numel() for t in ts])
        with torch.cuda.device(dev):
            rc = _lib.lib().sgb_nan_to_num_multi(ptrs, nums, len(ts), float(nan), float(posinf), float(neginf), _lib.stream_ptr(dev))
        _lib.check(rc, 'nan_to_num_multi')
    return tensors


class FlatGradAllReduce:
    """Average the gradients of a parameter list across ranks with a single flat all-reduce."""

    def __init__(self, params, group=None):
        self.params = [p for p in params]
        self.group = group

    def __call__(self, grads=None):
        """grads: explicit list of gradient tensors (the static ones of a captured graph); default = the .grad fields"""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        gs = [g for g in grads if g is not None] if grads is not None else [p.grad for p in self.params if p.grad is not None]
        if not gs:
            return 0
        flat = torch.cat([g.reshape(-1).to(torch.float32) for g in gs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        off = 0
        for g in gs:
            n = g.numel()
            g.copy_(flat[off:off + n].reshape(g.shape))
            off += n
        return flat.numel()


class FlatGrads:
    """The gradients of a parameter list gathered into ONE flat fp32 buffer, so that the per-phase bookkeeping of the reference
    loop (trainers.py:745-748,887-893) is a handful of launches that are all legal inside a CUDA graph: one multi-tensor copy
    of the gradients autograd produced into the buffer, one NCCL all-reduce of the buffer (what DistributedDataParallel's
    bucketed all-reduce amounts to), one scale, one nan_to_num -- instead of a torch.cat, an all-reduce between two graphs and
    ~200 per-tensor copies.  After `gather()` every `.grad` is a view of the buffer, which is what the optimizer reads.
    (Pointing `.grad` at the views BEFORE backward would save the copy but makes autograd accumulate in place: one extra
    elementwise add per parameter, ~200 launches per phase, measured slower.)  Parameters that receive no gradient in a
    phase keep `grad = None`, so Adam skips them exactly as in the reference."""

    def __init__(self, params, group=None):
        self.params = [p for p in params]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else None
        self.flat = torch.zeros([n], dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            assert p.dtype == torch.float32
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()

    def gather(self):
        """copy the gradients autograd left in .grad into the flat buffer (one multi-tensor launch) and re-point .grad at the views"""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if len(src) < len(self.params):
            self.flat.zero_()                      # slots of parameters without a gradient must not carry stale values into the sum
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                p.grad = v

    def all_reduce_mean(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(self.group)
        if world == 1:
            return 0
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)
        return self.flat.numel()

    def nan_to_num_(self):
        nan_to_num_([self.flat], nan=0, posinf=1e5, neginf=-1e5)


class Trainer:
    """One rank's replica: networks, optimisers, phases.  `iteration(real_u8)` is one reference training
    iteration (all phases due at this batch index) and returns a dict of scalar losses (device tensors)."""

    def __init__(self, cfg, device, rank=0, world_size=1):
        self.cfg, self.device, self.rank, self.world_size = cfg, torch.device(device), rank, world_size
        torch.manual_seed(cfg.seed * world_size + rank)       # trainers.py:507-508
        self.G, self.D = build_networks(cfg, self.device)
        if world_size > 1:      # same initial state on every rank: DistributedDataParallel broadcasts parameters AND buffers
            for t in (list(self.G.parameters()) + list(self.G.buffers()) +      # (noise_const, w_avg, ...) from rank 0 at construction
                      list(self.D.parameters()) + list(self.D.buffers())):
                dist.broadcast(t.data, src=0)
        self.G_ema = copy.deepcopy(self.G).eval().requires_grad_(False) if cfg.use_ema else None
        self.G.train().requires_grad_(False)
        self.D.train().requires_grad_(False)
        self.pl_mean = torch.zeros([], device=self.device)
        self.batch_idx = 0
        self.phases = []
        for name, module, interval, has_reg in [('G', self.G, cfg.g_reg_interval, cfg.use_ppl),
                                                ('D', self.D, cfg.d_reg_interval, cfg.use_r1)]:
            params = list(module.parameters())
            if not has_reg:
                opt = torch.optim.Adam(params, lr=cfg.lr, betas=tuple(cfg.betas), eps=1e-8, fused=True, capturable=cfg.cuda_graphs)
                self.phases.append(dict(name=name + 'main', module=module, opt=opt, interval=1))
            else:
                r = interval / (interval + 1)
                opt = torch.optim.Adam(params, lr=cfg.lr * r, betas=tuple(b ** r for b in cfg.betas), eps=1e-8, fused=True,
                                       capturable=cfg.cuda_graphs)
                self.phases.append(dict(name=name + 'main', module=module, opt=opt, interval=1))
                self.phases.append(dict(name=name + 'reg', module=module, opt=opt, interval=interval))
        # flat gradient buffers only where there is an all-reduce to feed: accumulating into pre-assigned .grad views costs one
        # extra elementwise add per parameter and phase (~200 launches), which a single GPU does not need
        self.use_flat = world_size > 1 or cfg.force_flat_grads
        self._flat = {id(self.G): FlatGrads(self.G.parameters()), id(self.D): FlatGrads(self.D.parameters())} if self.use_flat else {}
        for ph in self.phases:
            ph['flat'] = self._flat.get(id(ph['module']))
        # CUDA-graph state (cfg.cuda_graphs): static inputs, one graph pair per phase, shared memory pool
        self._graphs = None
        self.replayed_launches = 0      # libsgb200 kernel launches executed through graph replays
        self.static_z = None            # tests: {phase name: z tensor} captured as a static graph input instead of torch.randn
        self.static_pl_noise = None     # tests: the path-length noise image [N/2,3,R,R] instead of torch.randn_like

    # ---- forward helpers (losses_base.py:131-156)
    def run_G(self, z, return_ws=False):
        ws = self.G.mapping(z, None)
        if self.cfg.style_mixing_prob > 0:
            cutoff = torch.empty([], dtype=torch.int64, device=ws.device).random_(1, ws.shape[1])
            cutoff = torch.where(torch.rand([], device=ws.device) < self.cfg.style_mixing_prob, cutoff,
                                 torch.full_like(cutoff, ws.shape[1]))
            ws2 = self.G.mapping(torch.randn_like(z), None, skip_w_avg_update=True)
            # ws[:, cutoff:] = ws2[:, cutoff:] without a host sync on `cutoff`
            idx = torch.arange(ws.shape[1], device=ws.device).reshape(1, -1, 1)
            ws = torch.where(idx >= cutoff, ws2, ws)
        img = self.G.synthesis(ws, noise_mode=self.cfg.noise_mode)
        return (img, ws) if return_ws else img

    def run_D(self, img):
        return self.D(img, None)

    # ---- phases; each leaves gradients in .grad of the phase's module
    def phase_Gmain(self, z, gain):
        logits = self.run_D(self.run_G(z))
        loss = torch.nn.functional.softplus(-logits).mean()
        loss.mul(gain).backward()
        return loss.detach()

    def phase_Greg(self, z, gain, pl_noise=None):
        cfg = self.cfg
        n = z.shape[0] // cfg.pl_batch_shrink
        img, ws = self.run_G(z[:n], return_ws=True)
        if pl_noise is None:
            pl_noise = torch.randn_like(img)
        pl_noise = pl_noise / np.sqrt(img.shape[2] * img.shape[3])
        with conv2d_gradfix.no_weight_gradients():
            pl_grads, = torch.autograd.grad(outputs=[(img * pl_noise).sum()], inputs=[ws], create_graph=True, only_inputs=True)
        pl_lengths = pl_grads.square().sum(2).mean(1).sqrt()
        pl_mean = self.pl_mean.lerp(pl_lengths.mean(), cfg.pl_decay)
        self.pl_mean.copy_(pl_mean.detach())
        loss = (pl_lengths - pl_mean).square() * cfg.pl_weight
        (img[:, 0, 0, 0] * 0 + loss).mean().mul(gain).backward()
        return loss.detach().mean()

    def phase_Dmain(self, z, real, gain):
        with torch.no_grad():
            fake = self.run_G(z)
        gen_logits = self.run_D(fake)
        real_logits = self.run_D(real.detach().requires_grad_(self.cfg.use_r1))   # losses_base.py:72
        loss = torch.nn.functional.softplus(-real_logits).mean() + torch.nn.functional.softplus(gen_logits).mean()
        loss.mul(gain).backward()
        return loss.detach()

    def phase_Dreg(self, real, gain):
        real = real.detach().requires_grad_(True)
        real_logits = self.run_D(real)
        with conv2d_gradfix.no_weight_gradients():
            r1_grads, = torch.autograd.grad(outputs=[real_logits.sum()], inputs=[real], create_graph=True, only_inputs=True)
        pen = r1_grads.square().sum([1, 2, 3])
        loss = pen * (self.cfg.r1_gamma / 2)
        (real_logits * 0 + loss.unsqueeze(1)).mean().mul(gain).backward()
        return loss.detach().mean()

    def _phase_grads(self, ph, real, z):
        """forward + backward of one phase: leaves the gradients in .grad, returns the loss value"""
        opt, module, flat = ph['opt'], ph['module'], ph['flat']
        opt.zero_grad(set_to_none=True)
        module.requires_grad_(True)
        gain = ph['interval']
        name = ph['name']
        with fused_conv.phase_mode(first_order=name in ('Gmain', 'Dmain')):
            if name == 'Gmain':
                val = self.phase_Gmain(z, gain)
            elif name == 'Greg':
                val = self.phase_Greg(z, gain, pl_noise=self.static_pl_noise)
            elif name == 'Dmain':
                val = self.phase_Dmain(z, real, gain)
            else:
                val = self.phase_Dreg(real, gain)
        module.requires_grad_(False)
        if flat is not None:
            flat.gather()
        return val

    def _phase_update(self, ph):
        if ph['flat'] is not None:
            ph['flat'].nan_to_num_()               # trainers.py:745-748, all gradients of the module in one launch
        else:
            nan_to_num_([p.grad for p in ph['module'].parameters() if p.grad is not None], nan=0, posinf=1e5, neginf=-1e5)
        ph['opt'].step()

    def run_phase(self, ph, real, z):
        val = self._phase_grads(ph, real, z)
        if ph['flat'] is not None:
            ph['flat'].all_reduce_mean()           # trainers.py:887-893 (DistributedDataParallel) as one flat all-reduce
        self._phase_update(ph)
        return val

    # ---- CUDA graphs: every phase is static-shaped, so its ~1000 launches are captured once and replayed ----
    def _real_from_u8(self, real_u8):
        if real_u8.is_floating_point():                         # tests: images already in [-1, 1]
            return real_u8
        return real_u8.to(torch.float32) / 127.5 - 1            # trainers.py:716

    def _ema_update(self):
        cfg = self.cfg
        ema_nimg = cfg.ema_kimg * 1000
        beta = 0.5 ** (cfg.batch_gpu * self.world_size / max(ema_nimg, 1e-8))
        with torch.no_grad():
            ps = list(self.G.parameters())
            pe = list(self.G_ema.parameters())
            torch._foreach_lerp_(pe, ps, 1 - beta)           # p_ema = p.lerp(p_ema, beta)
            for b_ema, b in zip(self.G_ema.buffers(), self.G.buffers()):
                b_ema.copy_(b)

    def _snapshot_state(self):
        mods = [m for m in (self.G, self.D, self.G_ema) if m is not None]
        opts = {id(ph['opt']): ph['opt'] for ph in self.phases}.values()
        return dict(mods=[{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods],
                    opts=[copy.deepcopy(o.state_dict()) for o in opts], pl_mean=self.pl_mean.clone(),
                    rng=torch.cuda.get_rng_state(self.device))

    def _restore_state(self, snap):
        mods = [m for m in (self.G, self.D, self.G_ema) if m is not None]
        opts = {id(ph['opt']): ph['opt'] for ph in self.phases}.values()
        with torch.no_grad():
            for m, sd in zip(mods, snap['mods']):
                for k, v in m.state_dict().items():
                    if not torch.equal(v, sd[k]):      # untouched tensors (FIR filters, noise_const) keep their version counter:
                        v.copy_(sd[k])                 # ops cache host-side facts about constant buffers by (address, version)
            for o, sd in zip(opts, snap['opts']):
                if sd['state']:
                    o.load_state_dict(sd)
                    continue
                # fresh optimizer: keep the (lazily created) state tensors, reset exp_avg / exp_avg_sq / step to zero
                for st_ in o.state.values():
                    for v in st_.values():
                        if torch.is_tensor(v):
                            v.zero_()
            self.pl_mean.copy_(snap['pl_mean'])
        torch.cuda.set_rng_state(snap['rng'], self.device)

    def _build_graphs(self, real_u8):
        from . import _lib
        cfg = self.cfg
        dev = self.device
        multi = self.world_size > 1
        st = dict(real_u8=torch.empty_like(real_u8), graphs={}, pool=None)
        st['real_u8'].copy_(real_u8)
        # eager warm-up on a side stream: lazy initialisation (optimizer state, kernel attributes, library handles).
        # The warm-up runs real optimizer steps, so the training state is snapshotted before and restored after it:
        # iteration 0 of a graph-replayed run starts from the same weights / Adam moments / pl_mean / w_avg as an eager run.
        snap = self._snapshot_state()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                real = self._real_from_u8(st['real_u8'])
                for ph in self.phases:
                    z = torch.randn([cfg.batch_gpu, cfg.z_dim], device=dev)
                    self.run_phase(ph, real, z)
                if self.G_ema is not None:
                    self._ema_update()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore_state(snap)
        pool = torch.cuda.graph_pool_handle()
        for ph in self.phases:
            g1 = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g1, pool=pool):
                real = self._real_from_u8(st['real_u8'])
                z = (self.static_z or {}).get(ph['name'])
                if z is None:
                    z = torch.randn([cfg.batch_gpu, cfg.z_dim], device=dev)
                val = self._phase_grads(ph, real, z)
                if multi and cfg.allreduce_in_graph:
                    ph['flat'].all_reduce_mean()       # NCCL all-reduce of the flat gradient buffer, captured in the phase graph
                if not multi or cfg.allreduce_in_graph:
                    self._phase_update(ph)
            grads = [p.grad for p in ph['module'].parameters()]      # this graph's static gradient tensors (views of the flat buffer)
            g2 = None
            if multi and not cfg.allreduce_in_graph:     # A/B: all-reduce launched eagerly between two graphs
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=pool):
                    self._phase_update(ph)
            st['graphs'][ph['name']] = (g1, g2, val, _lib.launch_count() - n0, grads)
        if self.G_ema is not None:
            ge = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ge, pool=pool):
                self._ema_update()
            st['ema'] = ge
        self._graphs = st

    def _graph_iteration(self, real_u8, force_all_phases):
        if self._graphs is None:
            self._build_graphs(real_u8)
        st = self._graphs
        st['real_u8'].copy_(real_u8, non_blocking=True)
        out = {}
        for ph in self.phases:
            if not force_all_phases and self.batch_idx % ph['interval'] != 0:
                continue
            g1, g2, val, launches, grads = st['graphs'][ph['name']]
            g1.replay()
            if g2 is not None:
                ph['flat'].all_reduce_mean()
                g2.replay()
            self.replayed_launches += launches
            out[ph['name']] = val
        if self.G_ema is not None:
            st['ema'].replay()
        self.batch_idx += 1
        return out

    def iteration(self, real_u8, force_all_phases=False, eager=False):
        """real_u8: uint8 [batch_gpu, C, R, R] already on the device.  With cfg.cuda_graphs (and not `eager`) the
        returned loss tensors are the graphs' static outputs (overwritten by the next iteration)."""
        cfg = self.cfg
        if cfg.cuda_graphs and not eager:
            return self._graph_iteration(real_u8, force_all_phases)
        real = self._real_from_u8(real_u8)
        out = {}
        zs = torch.randn([len(self.phases), cfg.batch_gpu, cfg.z_dim], device=self.device)
        for ph, z in zip(self.phases, zs):
            if not force_all_phases and self.batch_idx % ph['interval'] != 0:
                continue
            out[ph['name']] = self.run_phase(ph, real, z)
        if self.G_ema is not None:
            self._ema_update()
        self.batch_idx += 1
        return out
