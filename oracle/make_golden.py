"""Generate tests/golden/*.npz from the REAL reference (imported from /root/reference, CPU,
impl='ref' path).  Runs only in the authoring container; the GPU box has no /root/reference, so
the vectors travel as committed fixtures.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden            # rewrites tests/golden/

What is pinned (the reference has no tests or vectors of its own):
  ops_bias_act.npz    bias_act ref  (stylegan2ada/torch_utils/ops/bias_act.py:93-123) fwd, dx, db, ddx
  ops_upfirdn2d.npz   upfirdn2d ref (upfirdn2d.py:168-208) + upsample2d/downsample2d/filter2d wrappers, fwd + dx
  ops_conv.npz        conv2d_resample (conv2d_resample.py:58-154) all branches, fwd + dx + dw
  ops_modconv.npz     modulated_conv2d (train_parts/generators.py:42-100) fused / non-fused, grads, 2nd order
  net_tiny.npz        train_parts G / D ('sg2_classic'), 32x32: state dicts, img, logits and the parameter
                      gradients of Gmain / Dmain / Dreg(R1) / Greg(PPL) computed by the reference's own
                      SG2Loss / R1reg / PPLreg (train_parts/losses_base.py, regularizations.py)
"""
import dataclasses
import json
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _import_reference():
    """Shims from SURVEY.md section 4: omegaconf stub + make_dataclass(eq=False)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
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


class Book:
    """npz writer: tensors under '<case>.<name>', parameters in a json manifest."""

    def __init__(self):
        self.arrays = {}
        self.manifest = []

    def add(self, params, **tensors):
        cid = f'c{len(self.manifest):03d}'
        self.manifest.append(dict(id=cid, **params))
        for k, v in tensors.items():
            if v is None:
                continue
            self.arrays[f'{cid}.{k}'] = v.detach().to(torch.float32).numpy() if isinstance(v, torch.Tensor) else np.asarray(v)

    def save(self, name):
        os.makedirs(OUT, exist_ok=True)
        path = os.path.join(OUT, name)
        np.savez_compressed(path, manifest=np.array(json.dumps(self.manifest)), **self.arrays)
        print(f'{name}: {len(self.manifest)} cases, {os.path.getsize(path) / 1024:.0f} KiB')


def gen_bias_act(ops):
    torch.manual_seed(0)
    bk = Book()
    shapes = [([2, 8, 5, 7], 1), ([3, 16], 1), ([2, 4, 6, 6], 1), ([4, 6, 3], 2)]
    for act in ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']:
        for (shape, dim) in shapes[:2] if act not in ('lrelu', 'linear') else shapes:
            for gain, clamp, alpha in [(None, None, None), (0.7071067811865476, 0.5, None), (1.0, 181.02, 0.1)]:
                x = (torch.randn(shape) * 2).requires_grad_(True)
                b = torch.randn(shape[dim]).requires_grad_(True)
                y = ops.bias_act.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref')
                dy = torch.randn_like(y)
                dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
                # second order: d/d(dy-direction) and d/dx of <dx, v>
                v = torch.randn_like(dx)
                dyl = dy.clone().requires_grad_(True)
                dx2, = torch.autograd.grad(
                    ops.bias_act.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref'),
                    [x], dyl, create_graph=True)
                g_dy, g_x = torch.autograd.grad((dx2 * v).sum(), [dyl, x], allow_unused=True)
                bk.add(dict(act=act, dim=dim, gain=gain, clamp=clamp, alpha=alpha),
                       x=x, b=b, y=y, dy=dy, dx=dx, db=db, v=v, g_dy=g_dy,
                       g_x=g_x if g_x is not None else torch.zeros_like(x))
    # no-bias case
    x = torch.randn(2, 3, 4, 4)
    bk.add(dict(act='lrelu', dim=1, gain=None, clamp=None, alpha=None, nobias=True),
           x=x, y=ops.bias_act.bias_act(x, None, act='lrelu', impl='ref'))
    bk.save('ops_bias_act.npz')


def gen_upfirdn2d(ops):
    torch.manual_seed(1)
    U = ops.upfirdn2d
    bk = Book()
    f4 = U.setup_filter([1, 3, 3, 1])
    f12 = U.setup_filter([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
                          0.4910559419267466, 0.787641141030194, 0.787641141030194, 0.4910559419267466,
                          -0.048311742585633, -0.11799011114819057, 0.0034907120842174702, 0.015404109327027373])
    f3x2 = torch.randn(2, 3)
    filters = {'f4': f4, 'f12': f12, 'f3x2': f3x2, 'none': None, 'f1d5': U.setup_filter([1, 4, 6, 4, 1], separable=True)}
    cases = [
        # (filter, shape, up, down, padding, flip, gain)  -- the four call-site forms of SURVEY 8(a4) first
        ('f4', [2, 3, 9, 9], 1, 1, [1, 1, 1, 1], False, 4),          # (i) post-convT FIR
        ('f4', [2, 3, 8, 8], 2, 1, [2, 1, 2, 1], False, 4),          # (ii) image upsample
        ('f4', [2, 4, 8, 8], 1, 1, [2, 2, 2, 2], False, 1),          # (iii) D pre-stride
        ('f4', [2, 4, 8, 8], 1, 2, [1, 1, 1, 1], False, 1),          # (iv) D skip decimate
        ('f4', [1, 2, 17, 13], 2, 1, [2, 1, 2, 1], True, 4),
        ('f4', [1, 2, 17, 13], 1, 2, [1, 1, 1, 1], True, 1),
        ('f4', [1, 2, 7, 5], [2, 1], [1, 2], [0, 3, 2, -1], False, 1.5),
        ('f3x2', [1, 2, 9, 10], 3, 2, [2, 1, 0, 3], False, 1),
        ('f3x2', [1, 2, 9, 10], 1, 1, [-1, 2, 1, -2], True, 2),
        ('f12', [1, 2, 20, 20], 2, 1, [5, 5, 5, 5], False, 4),
        ('f12', [1, 2, 24, 24], 1, 2, [5, 5, 5, 5], False, 1),
        ('f1d5', [1, 3, 11, 12], 2, 3, [2, 2, 3, 1], True, 1),
        ('none', [1, 2, 5, 6], 2, 1, 0, False, 1),
        ('f4', [1, 1, 4, 4], 1, 1, 0, False, 1),
    ]
    for fname, shape, up, down, pad, flip, gain in cases:
        f = filters[fname]
        x = torch.randn(shape).requires_grad_(True)
        y = U.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain, impl='ref')
        dy = torch.randn_like(y)
        dx, = torch.autograd.grad(y, [x], dy)
        bk.add(dict(fn='upfirdn2d', f=fname, up=up, down=down, padding=pad, flip_filter=flip, gain=gain),
               x=x, f=f, y=y, dy=dy, dx=dx)
    for fn in ['upsample2d', 'downsample2d', 'filter2d']:
        for fname, shape in [('f4', [2, 3, 8, 8]), ('f12', [1, 2, 16, 16])]:
            f = filters[fname]
            x = torch.randn(shape)
            y = getattr(U, fn)(x, f, impl='ref')
            bk.add(dict(fn=fn, f=fname), x=x, f=f, y=y)
    bk.save('ops_upfirdn2d.npz')


def gen_conv(ops):
    torch.manual_seed(2)
    bk = Book()
    f4 = ops.upfirdn2d.setup_filter([1, 3, 3, 1])
    cases = [
        # (N, Ci, Co, H, W, k, up, down, padding, groups, flip_weight)
        (2, 4, 6, 8, 8, 3, 1, 1, 1, 1, True),
        (2, 4, 6, 8, 8, 3, 1, 1, 1, 1, False),
        (2, 4, 6, 8, 8, 3, 2, 1, 1, 1, False),     # G up layer
        (2, 4, 6, 8, 8, 3, 2, 1, 1, 1, True),
        (2, 4, 6, 8, 8, 3, 1, 2, 1, 1, True),      # D conv1
        (2, 4, 6, 8, 8, 1, 1, 2, 0, 1, True),      # D skip
        (2, 4, 6, 8, 8, 1, 2, 1, 0, 1, False),     # G resnet skip
        (2, 4, 3, 8, 8, 1, 1, 1, 0, 1, True),      # toRGB
        (2, 3, 8, 8, 8, 1, 1, 1, 0, 1, True),      # fromRGB
        (1, 6, 8, 7, 9, 3, 1, 1, 1, 2, True),      # grouped
        (1, 6, 8, 6, 6, 3, 2, 1, 1, 2, False),     # grouped + up
        (1, 4, 4, 8, 8, 3, 2, 2, 1, 1, True),      # up and down
        (1, 4, 4, 8, 8, 3, 1, 1, [1, 0, 2, 1], 1, True),   # generic fallback (asymmetric pad)
        (1, 4, 4, 5, 5, 3, 1, 1, 0, 1, True),      # pad 0
    ]
    for (n, ci, co, h, w_, k, up, down, pad, groups, fw) in cases:
        x = torch.randn(n, ci, h, w_).requires_grad_(True)
        w = (torch.randn(co, ci // groups, k, k) / np.sqrt(ci * k * k)).requires_grad_(True)
        y = ops.conv2d_resample.conv2d_resample(x, w, f=f4, up=up, down=down, padding=pad, groups=groups, flip_weight=fw)
        dy = torch.randn_like(y)
        dx, dw = torch.autograd.grad(y, [x, w], dy)
        bk.add(dict(up=up, down=down, padding=pad, groups=groups, flip_weight=fw),
               x=x, w=w, f=f4, y=y, dy=dy, dx=dx, dw=dw)
    bk.save('ops_conv.npz')


def gen_modconv(ops, gen_mod):
    torch.manual_seed(3)
    bk = Book()
    f4 = ops.upfirdn2d.setup_filter([1, 3, 3, 1])
    mc = gen_mod.modulated_conv2d
    for (n, ci, co, h, k, up, demod, fused, use_noise, dtype) in [
        (2, 8, 6, 8, 3, 1, True, False, True, 'f32'),
        (2, 8, 6, 8, 3, 1, True, True, True, 'f32'),
        (2, 8, 6, 8, 3, 2, True, False, True, 'f32'),
        (2, 8, 6, 8, 3, 2, True, True, False, 'f32'),
        (2, 8, 3, 8, 1, 1, False, False, False, 'f32'),
        (2, 8, 3, 8, 1, 1, False, True, False, 'f32'),
        (2, 8, 6, 8, 3, 1, False, False, True, 'f32'),
        (3, 16, 16, 4, 3, 1, True, False, True, 'f32'),
    ]:
        x = torch.randn(n, ci, h, h).requires_grad_(True)
        w = torch.randn(co, ci, k, k).requires_grad_(True)
        s = (torch.randn(n, ci) + 1).requires_grad_(True)
        ho = h * up
        noise = (torch.randn(n, 1, ho, ho) * 0.3) if use_noise else None
        pad = k // 2
        y = mc(x=x, weight=w, styles=s, noise=noise, up=up, padding=pad, resample_filter=f4, demodulate=demod,
               flip_weight=(up == 1), fused_modconv=fused)
        dy = torch.randn_like(y)
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy, create_graph=True)
        # PPL-like second-order term: gradient of |ds|^2 wrt w and s
        g2w, g2s = torch.autograd.grad(ds.square().sum(), [w, s], allow_unused=True)
        bk.add(dict(up=up, padding=pad, demodulate=demod, fused_modconv=fused, flip_weight=(up == 1), dtype=dtype),
               x=x, w=w, s=s, noise=noise, f=f4, y=y, dy=dy, dx=dx, dw=dw, ds=ds, g2w=g2w, g2s=g2s)
    bk.save('ops_modconv.npz')


def _to_easy(obj, dnnlib):
    if dataclasses.is_dataclass(obj):
        return dnnlib.EasyDict({f.name: _to_easy(getattr(obj, f.name), dnnlib) for f in dataclasses.fields(obj)})
    if isinstance(obj, dict):
        return dnnlib.EasyDict({k: _to_easy(v, dnnlib) for k, v in obj.items()})
    return obj


TINY = dict(img_resolution=32, z_dim=64, w_dim=64, channel_base=512, channel_max=32, map_layers=2,
            mbstd_group_size=2, d_arch='resnet')


def build_reference_nets(cfg=TINY, seed=0, noise_strength=0.1, num_fp16_res=0, conv_clamp=None):
    """train_parts 'sg2_classic' G and D built as SURVEY.md section 4 describes."""
    _import_reference()
    import stylegan2ada.dnnlib as dnnlib
    from train_parts.generators import generators
    from train_parts.discriminators import discriminators
    torch.manual_seed(seed)
    gk = _to_easy(generators.args['sg2_classic'](), dnnlib)
    gk.update(z_dim=cfg['z_dim'], w_dim=cfg['w_dim'], c_dim=0, img_resolution=cfg['img_resolution'], img_channels=3)
    gk.mapping_kwargs.num_layers = cfg['map_layers']
    gk.synthesis_kwargs.channel_base = cfg['channel_base']
    gk.synthesis_kwargs.channel_max = cfg['channel_max']
    gk.synthesis_kwargs.num_fp16_res = num_fp16_res
    gk.synthesis_kwargs.block_kwargs.conv_clamp = conv_clamp
    G = generators['sg2_classic'](**gk)
    dk = _to_easy(discriminators.args['sg2_classic'](), dnnlib)
    dk.update(c_dim=0, img_resolution=cfg['img_resolution'], img_channels=3, architecture=cfg['d_arch'],
              channel_base=cfg['channel_base'], channel_max=cfg['channel_max'], num_fp16_res=num_fp16_res,
              conv_clamp=conv_clamp)
    dk.epilogue_kwargs.mbstd_group_size = cfg['mbstd_group_size']
    D = discriminators['sg2_classic'](**dk)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith('noise_strength'):
                p.fill_(noise_strength)
            if name.endswith('.bias') and 'affine' not in name:
                p.copy_(torch.randn_like(p) * 0.1)
        for name, p in D.named_parameters():
            if name.endswith('.bias'):
                p.copy_(torch.randn_like(p) * 0.1)
    return G, D


class _ConstNoiseSynthesis(torch.nn.Module):
    def __init__(self, syn):
        super().__init__()
        self.syn = syn

    def forward(self, ws):
        return self.syn(ws, noise_mode='const')


def gen_net():
    _import_reference()
    from train_parts.losses_base import losses_arch
    G, D = build_reference_nets()
    G.train()
    D.train()
    arrays = {}
    for k, v in G.state_dict().items():
        arrays['G.' + k] = v.numpy()
    for k, v in D.state_dict().items():
        arrays['D.' + k] = v.numpy()
    n = 4
    torch.manual_seed(10)
    z = torch.randn(n, TINY['z_dim'])
    c = torch.zeros(n, 0)
    real = torch.rand(n, 3, 32, 32) * 2 - 1
    arrays['z'] = z.numpy()
    arrays['real'] = real.numpy()

    with torch.no_grad():
        ws = G.mapping(z, c, skip_w_avg_update=True)
        img = G.synthesis(ws, noise_mode='const')
        logits = D(img, c)
    arrays['ws'] = ws.numpy()
    arrays['img'] = img.numpy()
    arrays['logits'] = logits.numpy()

    loss = losses_arch['sg2'](G_mapping=G.mapping, G_synthesis=_ConstNoiseSynthesis(G.synthesis), D=D, device='cpu',
                              gen_regs=[('ppl', dict(pl_batch_shrink=2, pl_decay=0.01, pl_weight=2.0))],
                              dis_regs=[('r1', dict(r1_gamma=10.0))], loss='softplus', style_mixing_prob=0)
    w_avg0 = G.mapping.w_avg.clone()

    def run(phase, gain):
        for p in list(G.parameters()) + list(D.parameters()):
            p.grad = None
        G.requires_grad_(phase.startswith('G'))
        D.requires_grad_(phase.startswith('D'))
        G.mapping.w_avg.copy_(w_avg0)
        torch.manual_seed(20)
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=gain)
        mod, tag = (G, 'G.') if phase.startswith('G') else (D, 'D.')
        for k, p in mod.named_parameters():
            if p.grad is not None:
                arrays[f'{phase}.grad.{tag}{k}'] = p.grad.numpy().copy()

    run('Gmain', 1)
    run('Dmain', 1)
    run('Dreg', 4)
    run('Greg', 16)
    # the PPL noise the reference drew (first draw after manual_seed(20); N/2 images)
    torch.manual_seed(20)
    arrays['pl_noise'] = torch.randn(n // 2, 3, 32, 32).numpy()
    arrays['meta'] = np.array(json.dumps(dict(cfg=TINY, noise_strength=0.1, n=n, r1_gamma=10.0, pl_weight=2.0,
                                              pl_decay=0.01, gains=dict(Gmain=1, Dmain=1, Dreg=4, Greg=16))))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, 'net_tiny.npz')
    np.savez_compressed(path, **arrays)
    print(f'net_tiny.npz: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


def main():
    _import_reference()
    from stylegan2ada.torch_utils import ops as _pkg  # noqa: F401
    from stylegan2ada.torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, fma
    import train_parts.generators as gen_mod
    ops = types.SimpleNamespace(bias_act=bias_act, upfirdn2d=upfirdn2d, conv2d_resample=conv2d_resample, fma=fma)
    gen_bias_act(ops)
    gen_upfirdn2d(ops)
    gen_conv(ops)
    gen_modconv(ops, gen_mod)
    gen_net()


if __name__ == '__main__':
    main()
