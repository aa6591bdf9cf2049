"""CPU oracle for the callers of the hot path: StyleGAN2 G / D forward and the four training
phases (Gmain, Dmain, Greg = path-length, Dreg = R1).  TEST INFRASTRUCTURE ONLY (see
``oracle/ref_ops.py`` for the usage rule).

Functional restatement: parameters live in a flat ``dict`` whose keys equal the reference
modules' ``state_dict()`` keys (e.g. ``synthesis.b8.conv0.affine.weight``), so a state dict
exported from the reference (``oracle/make_golden.py``) drives this code unchanged.  Pinned
against ``tests/golden/net_*.npz`` by ``tests/test_oracle_golden.py``.

Follows (paths relative to /root/reference):
  FullyConnectedLayer   train_parts/generators.py:104-134
  MappingNetwork        train_parts/generators.py:190-269
  SynthesisLayer        train_parts/generators.py:272-329
  ToRGBLayer            train_parts/generators.py:333-348
  SynthesisBlock        train_parts/generators.py:354-458   (architecture 'skip', the default)
  SynthesisNetwork      train_parts/generators.py:464-519
  Conv2dLayer           train_parts/discriminators.py:78-124
  DiscriminatorBlock    train_parts/discriminators.py:215-302
  MinibatchStdLayer     train_parts/discriminators.py:306-328
  DiscriminatorEpilogue train_parts/discriminators.py:332-389
  phases                train_parts/losses_base.py:43-109, regularizations.py:11-56, losses.py:47-58
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import ref_ops as R


@dataclass
class NetConfig:
    img_resolution: int = 64
    img_channels: int = 3
    z_dim: int = 512
    w_dim: int = 512
    channel_base: int = 32768
    channel_max: int = 512
    map_layers: int = 8
    num_fp16_res: int = 0
    conv_clamp: float = None
    d_arch: str = 'resnet'
    mbstd_group_size: int = 4
    mbstd_num_channels: int = 1

    @property
    def res_log2(self):
        return int(round(math.log2(self.img_resolution)))

    @property
    def g_resolutions(self):
        return [2 ** i for i in range(2, self.res_log2 + 1)]

    @property
    def d_resolutions(self):
        return [2 ** i for i in range(self.res_log2, 2, -1)]

    def channels(self, res):
        return min(self.channel_base // res, self.channel_max)

    @property
    def fp16_resolution(self):
        return max(2 ** (self.res_log2 + 1 - self.num_fp16_res), 8)

    @property
    def num_ws(self):
        # one per conv layer plus the last block's toRGB (generators.py:489-500)
        return 2 * len(self.g_resolutions)


# ----------------------------------------------------------------------------------------
# parameter construction (same shapes / init distributions as the reference modules)

def g_param_shapes(cfg):
    shp = {}
    feats = [cfg.z_dim] + [cfg.w_dim] * cfg.map_layers
    for i in range(cfg.map_layers):
        shp[f'mapping.fc{i}.weight'] = (feats[i + 1], feats[i])
        shp[f'mapping.fc{i}.bias'] = (feats[i + 1],)
    shp['mapping.w_avg'] = (cfg.w_dim,)
    for res in cfg.g_resolutions:
        co = cfg.channels(res)
        ci = cfg.channels(res // 2) if res > 4 else 0
        pre = f'synthesis.b{res}.'
        if ci == 0:
            shp[pre + 'const'] = (co, res, res)
        layers = (['conv0'] if ci else []) + ['conv1']
        for name in layers:
            cin = ci if name == 'conv0' else co
            shp[pre + name + '.affine.weight'] = (cin, cfg.w_dim)
            shp[pre + name + '.affine.bias'] = (cin,)
            shp[pre + name + '.weight'] = (co, cin, 3, 3)
            shp[pre + name + '.noise_const'] = (res, res)
            shp[pre + name + '.noise_strength'] = ()
            shp[pre + name + '.bias'] = (co,)
        shp[pre + 'torgb.affine.weight'] = (co, cfg.w_dim)
        shp[pre + 'torgb.affine.bias'] = (co,)
        shp[pre + 'torgb.weight'] = (cfg.img_channels, co, 1, 1)
        shp[pre + 'torgb.bias'] = (cfg.img_channels,)
    return shp


def d_param_shapes(cfg):
    shp = {}
    for res in cfg.d_resolutions:
        tmp = cfg.channels(res)
        out = cfg.channels(res // 2)
        pre = f'b{res}.'
        if res == cfg.img_resolution or cfg.d_arch == 'skip':
            shp[pre + 'fromrgb.weight'] = (tmp, cfg.img_channels, 1, 1)
            shp[pre + 'fromrgb.bias'] = (tmp,)
        shp[pre + 'conv0.weight'] = (tmp, tmp, 3, 3)
        shp[pre + 'conv0.bias'] = (tmp,)
        shp[pre + 'conv1.weight'] = (out, tmp, 3, 3)
        shp[pre + 'conv1.bias'] = (out,)
        if cfg.d_arch == 'resnet':
            shp[pre + 'skip.weight'] = (out, tmp, 1, 1)
    c4 = cfg.channels(4)
    if cfg.d_arch == 'skip':
        shp['b4.fromrgb.weight'] = (c4, cfg.img_channels, 1, 1)
        shp['b4.fromrgb.bias'] = (c4,)
    shp['b4.conv.weight'] = (c4, c4 + cfg.mbstd_num_channels, 3, 3)
    shp['b4.conv.bias'] = (c4,)
    shp['b4.fc.weight'] = (c4, c4 * 16)
    shp['b4.fc.bias'] = (c4,)
    shp['b4.out.weight'] = (1, c4)
    shp['b4.out.bias'] = (1,)
    return shp


G_BUFFERS = ('noise_const', 'w_avg')


def init_params(shapes, seed, noise_strength=0.0, dtype=torch.float32):
    """randn weights, zero biases, affine bias 1, mapping weights / lr_multiplier (0.01)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, s in shapes.items():
        if k.endswith('noise_strength'):
            v = torch.full(s, float(noise_strength))
        elif k.endswith('w_avg'):
            v = torch.zeros(s)
        elif k.endswith('affine.bias'):
            v = torch.ones(s)
        elif k.endswith('.bias'):
            v = torch.zeros(s)
        elif k.startswith('mapping.') and k.endswith('.weight'):
            v = torch.randn(s, generator=g) / 0.01
        else:
            v = torch.randn(s, generator=g)
        out[k] = v.to(dtype)
    return out


# ----------------------------------------------------------------------------------------
# layers

def _fc(x, w, b, act='linear', lr_mult=1.0):
    wg = w.to(x.dtype) * (lr_mult / math.sqrt(w.shape[1]))
    if b is not None:
        b = b.to(x.dtype)
        if lr_mult != 1:
            b = b * lr_mult
    if act == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, wg.t())
    return R.bias_act(x.matmul(wg.t()), b, act=act)


def _base(t):
    """fp32 (the reference's type), or fp64 when the oracle is run in double to serve as an exact reference"""
    return torch.float64 if t.dtype == torch.float64 else torch.float32


def g_mapping(P, z, cfg):
    x = z.to(_base(z))
    x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    for i in range(cfg.map_layers):
        x = _fc(x, P[f'mapping.fc{i}.weight'], P[f'mapping.fc{i}.bias'], act='lrelu', lr_mult=0.01)
    return x.unsqueeze(1).repeat(1, cfg.num_ws, 1)


_FILT = None


def _filt():
    global _FILT
    if _FILT is None:
        _FILT = R.setup_filter([1, 3, 3, 1])
    return _FILT


def _synth_layer(P, pre, x, w, cfg, up, res, noise, gain=1.0, fused_modconv=False):
    styles = _fc(w, P[pre + 'affine.weight'], P[pre + 'affine.bias'])
    nz = None
    if noise == 'const':
        nz = P[pre + 'noise_const'] * P[pre + 'noise_strength']
    elif isinstance(noise, dict):
        nz = noise[pre] * P[pre + 'noise_strength']
    x = R.modulated_conv2d(x, P[pre + 'weight'], styles, noise=nz, up=up, padding=1, resample_filter=_filt(),
                           flip_weight=(up == 1), fused_modconv=fused_modconv)
    clamp = cfg.conv_clamp * gain if cfg.conv_clamp is not None else None
    return R.bias_act(x, P[pre + 'bias'].to(x.dtype), act='lrelu', gain=math.sqrt(2.0) * gain, clamp=clamp)


def _torgb(P, pre, x, w, cfg, fused_modconv=False):
    ci = P[pre + 'weight'].shape[1]
    styles = _fc(w, P[pre + 'affine.weight'], P[pre + 'affine.bias']) * (1 / math.sqrt(ci))
    x = R.modulated_conv2d(x, P[pre + 'weight'], styles, demodulate=False, fused_modconv=fused_modconv)
    return R.bias_act(x, P[pre + 'bias'].to(x.dtype), clamp=cfg.conv_clamp)


def g_synthesis(P, ws, cfg, noise='const', fused_modconv=False):
    """noise: 'const' | 'none' | dict(layer prefix -> [N,1,R,R] tensor)."""
    ws = ws.to(_base(ws))
    x = img = None
    widx = 0
    for res in cfg.g_resolutions:
        pre = f'synthesis.b{res}.'
        dtype = torch.float16 if res >= cfg.fp16_resolution else ws.dtype       # ws.dtype: fp32, or fp64 for a double oracle run
        if res == 4:
            x = P[pre + 'const'].to(dtype).unsqueeze(0).repeat(ws.shape[0], 1, 1, 1)
            x = _synth_layer(P, pre + 'conv1.', x, ws[:, widx], cfg, 1, res, noise, fused_modconv=fused_modconv)
            nconv = 1
        else:
            x = x.to(dtype)
            x = _synth_layer(P, pre + 'conv0.', x, ws[:, widx], cfg, 2, res, noise, fused_modconv=fused_modconv)
            x = _synth_layer(P, pre + 'conv1.', x, ws[:, widx + 1], cfg, 1, res, noise, fused_modconv=fused_modconv)
            nconv = 2
        if img is not None:
            img = R.upsample2d(img, _filt())
        y = _torgb(P, pre + 'torgb.', x, ws[:, widx + nconv], cfg, fused_modconv=fused_modconv).to(ws.dtype)
        img = y if img is None else img + y
        widx += nconv
    return img


def _conv_layer(P, pre, x, act='linear', up=1, down=1, gain=1.0, clamp=None):
    w = P[pre + 'weight']
    k = w.shape[-1]
    wg = (w * (1 / math.sqrt(w.shape[1] * k * k))).to(x.dtype)
    b = P.get(pre + 'bias')
    b = b.to(x.dtype) if b is not None else None
    x = R.conv2d_resample(x, wg, f=_filt(), up=up, down=down, padding=k // 2, flip_weight=(up == 1))
    _, g0 = R.act_defaults(act)
    c = clamp * gain if clamp is not None else None
    return R.bias_act(x, b, act=act, gain=g0 * gain, clamp=c)


def _mbstd(x, group_size, nch):
    n, c, h, w = x.shape
    g = min(group_size, n) if group_size is not None else n
    y = x.reshape(g, -1, nch, c // nch, h, w)
    y = y - y.mean(dim=0)
    y = (y.square().mean(dim=0) + 1e-8).sqrt()
    y = y.mean(dim=[2, 3, 4]).reshape(-1, nch, 1, 1).repeat(g, 1, h, w)
    return torch.cat([x, y], dim=1)


def d_forward(P, img, cfg):
    base = _base(img)
    x = None
    for res in cfg.d_resolutions:
        pre = f'b{res}.'
        dtype = torch.float16 if res >= cfg.fp16_resolution else base
        if x is not None:
            x = x.to(dtype)
        if res == cfg.img_resolution or cfg.d_arch == 'skip':
            y = _conv_layer(P, pre + 'fromrgb.', img.to(dtype), act='lrelu', clamp=cfg.conv_clamp)
            x = y if x is None else x + y
            img = R.downsample2d(img.to(dtype), _filt()) if cfg.d_arch == 'skip' else None
        if cfg.d_arch == 'resnet':
            y = _conv_layer(P, pre + 'skip.', x, down=2, gain=math.sqrt(0.5))
            x = _conv_layer(P, pre + 'conv0.', x, act='lrelu', clamp=cfg.conv_clamp)
            x = _conv_layer(P, pre + 'conv1.', x, act='lrelu', down=2, gain=math.sqrt(0.5), clamp=cfg.conv_clamp)
            x = y + x
        else:
            x = _conv_layer(P, pre + 'conv0.', x, act='lrelu', clamp=cfg.conv_clamp)
            x = _conv_layer(P, pre + 'conv1.', x, act='lrelu', down=2, clamp=cfg.conv_clamp)
    x = x.to(base)
    if cfg.d_arch == 'skip':
        x = x + _conv_layer(P, 'b4.fromrgb.', img.to(base), act='lrelu')
    if cfg.mbstd_num_channels > 0:
        x = _mbstd(x, cfg.mbstd_group_size, cfg.mbstd_num_channels)
    x = _conv_layer(P, 'b4.conv.', x, act='lrelu', clamp=cfg.conv_clamp)
    x = _fc(x.flatten(1), P['b4.fc.weight'], P['b4.fc.bias'], act='lrelu')
    return _fc(x, P['b4.out.weight'], P['b4.out.bias'])


# ----------------------------------------------------------------------------------------
# training phases: each returns (scalar loss value, dict of parameter gradients)

def _trainable(P, buffers=G_BUFFERS):
    return {k: v for k, v in P.items() if not k.endswith(buffers)}


def _grads(loss, params):
    keys = list(params)
    gs = torch.autograd.grad(loss, [params[k] for k in keys], allow_unused=True)
    return {k: g for k, g in zip(keys, gs) if g is not None}


def _leafify(P):
    out = {}
    for k, v in P.items():
        if k.endswith(G_BUFFERS):
            out[k] = v
        else:
            out[k] = v.detach().clone().requires_grad_(True)
    return out


def phase_gmain(GP, DP, z, gcfg, dcfg, noise='const', gain=1.0):
    GP = _leafify(GP)
    img = g_synthesis(GP, g_mapping(GP, z, gcfg), gcfg, noise=noise)
    logits = d_forward(DP, img, dcfg)
    loss = torch.nn.functional.softplus(-logits).mean() * gain
    return loss.detach(), _grads(loss, _trainable(GP)), img.detach()


def phase_dmain(GP, DP, z, real, gcfg, dcfg, noise='const', gain=1.0):
    DP = _leafify(DP)
    with torch.no_grad():
        fake = g_synthesis(GP, g_mapping(GP, z, gcfg), gcfg, noise=noise)
    lf = d_forward(DP, fake, dcfg)
    lr = d_forward(DP, real, dcfg)
    loss = (torch.nn.functional.softplus(-lr).mean() + torch.nn.functional.softplus(lf).mean()) * gain
    return loss.detach(), _grads(loss, DP)


def phase_dreg(DP, real, dcfg, r1_gamma=10.0, gain=1.0):
    DP = _leafify(DP)
    real = real.detach().requires_grad_(True)
    logits = d_forward(DP, real, dcfg)
    g, = torch.autograd.grad(logits.sum(), [real], create_graph=True)
    pen = g.square().sum([1, 2, 3])
    loss = (logits * 0 + (pen * (r1_gamma / 2)).unsqueeze(1)).mean() * gain
    return loss.detach(), _grads(loss, DP), pen.detach()


def phase_greg(GP, z, pl_noise, gcfg, pl_mean=0.0, pl_decay=0.01, pl_weight=2.0, noise='const', gain=1.0):
    """pl_noise: [N,C,R,R] standard normal; divided by sqrt(R*R) here as regularizations.py:26."""
    GP = _leafify(GP)
    ws = g_mapping(GP, z, gcfg)
    img = g_synthesis(GP, ws, gcfg, noise=noise)
    pn = pl_noise / np.sqrt(img.shape[2] * img.shape[3])
    g, = torch.autograd.grad((img * pn).sum(), [ws], create_graph=True)
    lengths = g.square().sum(2).mean(1).sqrt()
    mean = pl_mean + pl_decay * (lengths.mean() - pl_mean)
    pen = (lengths - mean).square()          # pl_mean is NOT detached in regularizations.py:31-33
    loss = (img[:, 0, 0, 0] * 0 + pen * pl_weight).mean() * gain
    return loss.detach(), _grads(loss, _trainable(GP)), lengths.detach()
