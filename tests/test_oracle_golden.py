"""CPU: the oracle (oracle/ref_ops.py, ref_networks.py) against the golden vectors that
oracle/make_golden.py produced by running the real reference (impl='ref')."""
import json

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import ref_networks as RN
from helpers import Golden, assert_close

TOL = 2e-5   # same arithmetic, different op order at most


def test_bias_act_golden():
    g = Golden('ops_bias_act.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x').requires_grad_(True)
        if c.get('nobias'):
            assert_close(R.bias_act(x, None, act=c['act']), g.t(cid, 'y'), TOL, cid)
            continue
        b = g.t(cid, 'b').requires_grad_(True)
        kw = dict(dim=c['dim'], act=c['act'], alpha=c['alpha'], gain=c['gain'], clamp=c['clamp'])
        y = R.bias_act(x, b, **kw)
        assert_close(y, g.t(cid, 'y'), TOL, f'{c} y')
        dy = g.t(cid, 'dy')
        dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
        assert_close(dx, g.t(cid, 'dx'), TOL, f'{c} dx')
        assert_close(db, g.t(cid, 'db'), TOL, f'{c} db')
        dyl = dy.clone().requires_grad_(True)
        dx2, = torch.autograd.grad(R.bias_act(x, b, **kw), [x], dyl, create_graph=True)
        g_dy, g_x = torch.autograd.grad((dx2 * g.t(cid, 'v')).sum(), [dyl, x], allow_unused=True)
        assert_close(g_dy, g.t(cid, 'g_dy'), TOL, f'{c} g_dy')
        assert_close(g_x if g_x is not None else torch.zeros_like(x), g.t(cid, 'g_x'), 1e-4, f'{c} g_x')


def test_upfirdn2d_golden():
    g = Golden('ops_upfirdn2d.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x').requires_grad_(True)
        f = g.t(cid, 'f')
        if c['fn'] == 'upfirdn2d':
            y = R.upfirdn2d(x, f, up=c['up'], down=c['down'], padding=c['padding'], flip_filter=c['flip_filter'], gain=c['gain'])
            assert_close(y, g.t(cid, 'y'), TOL, f'{c} y')
            dx, = torch.autograd.grad(y, [x], g.t(cid, 'dy'))
            assert_close(dx, g.t(cid, 'dx'), TOL, f'{c} dx')
        else:
            y = getattr(R, c['fn'])(x, f)
            assert_close(y, g.t(cid, 'y'), TOL, f'{c} y')


def test_setup_filter_matches_reference_filter():
    g = Golden('ops_upfirdn2d.npz')
    f = g.t('c000', 'f')
    assert_close(R.setup_filter([1, 3, 3, 1]), f, 1e-7, 'setup_filter')
    assert R.setup_filter([1] * 8).ndim == 1        # >= 8 taps stay separable (upfirdn2d.py:103-106)


def test_conv2d_resample_golden():
    g = Golden('ops_conv.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x').requires_grad_(True)
        w = g.t(cid, 'w').requires_grad_(True)
        y = R.conv2d_resample(x, w, f=g.t(cid, 'f'), up=c['up'], down=c['down'], padding=c['padding'],
                              groups=c['groups'], flip_weight=c['flip_weight'])
        assert_close(y, g.t(cid, 'y'), TOL, f'{c} y')
        dx, dw = torch.autograd.grad(y, [x, w], g.t(cid, 'dy'))
        assert_close(dx, g.t(cid, 'dx'), TOL, f'{c} dx')
        assert_close(dw, g.t(cid, 'dw'), TOL, f'{c} dw')


def test_modulated_conv2d_golden():
    g = Golden('ops_modconv.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x').requires_grad_(True)
        w = g.t(cid, 'w').requires_grad_(True)
        s = g.t(cid, 's').requires_grad_(True)
        y = R.modulated_conv2d(x, w, s, noise=g.t(cid, 'noise'), up=c['up'], padding=c['padding'],
                               resample_filter=g.t(cid, 'f'), demodulate=c['demodulate'],
                               flip_weight=c['flip_weight'], fused_modconv=c['fused_modconv'])
        assert_close(y, g.t(cid, 'y'), TOL, f'{c} y')
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], g.t(cid, 'dy'), create_graph=True)
        assert_close(dx, g.t(cid, 'dx'), TOL, f'{c} dx')
        assert_close(dw, g.t(cid, 'dw'), TOL, f'{c} dw')
        assert_close(ds, g.t(cid, 'ds'), TOL, f'{c} ds')
        g2w, g2s = torch.autograd.grad(ds.square().sum(), [w, s], allow_unused=True)
        for got, key in ((g2w, 'g2w'), (g2s, 'g2s')):
            want = g.t(cid, key)
            assert (got is None) == (want is None), f'{c} {key} presence'
            if want is not None:
                assert_close(got, want, 1e-4, f'{c} {key}')


@pytest.fixture(scope='module')
def net():
    z = np.load(__import__('os').path.join(__import__('helpers').GOLDEN, 'net_tiny.npz'))
    meta = json.loads(str(z['meta']))
    cfg = RN.NetConfig(**meta['cfg'])
    GP = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('G.')}
    DP = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('D.')}
    return z, meta, cfg, GP, DP


def test_param_shapes_match_reference_state_dict(net):
    z, meta, cfg, GP, DP = net
    gs = RN.g_param_shapes(cfg)
    ds = RN.d_param_shapes(cfg)
    gref = {k: tuple(v.shape) for k, v in GP.items() if not k.endswith('resample_filter')}
    dref = {k: tuple(v.shape) for k, v in DP.items() if not k.endswith('resample_filter')}
    assert gs == gref
    assert ds == dref


def test_network_forward_golden(net):
    z, meta, cfg, GP, DP = net
    zz = torch.from_numpy(z['z'])
    with torch.no_grad():
        ws = RN.g_mapping(GP, zz, cfg)
        img = RN.g_synthesis(GP, ws, cfg, noise='const')
        logits = RN.d_forward(DP, img, cfg)
        img_fused = RN.g_synthesis(GP, ws, cfg, noise='const', fused_modconv=True)
    assert_close(ws, torch.from_numpy(z['ws']), TOL, 'ws')
    assert_close(img, torch.from_numpy(z['img']), TOL, 'img')
    assert_close(img_fused, torch.from_numpy(z['img']), 1e-4, 'img fused')
    assert_close(logits, torch.from_numpy(z['logits']), TOL, 'logits')


def _check_grads(z, phase, tag, grads, tol):
    keys = [k for k in z.files if k.startswith(f'{phase}.grad.{tag}')]
    assert keys
    for k in keys:
        name = k[len(f'{phase}.grad.{tag}'):]
        assert name in grads, f'{phase}: missing grad for {name}'
        assert_close(grads[name], torch.from_numpy(z[k]), tol, f'{phase} {name}')


def test_phase_gmain_golden(net):
    z, meta, cfg, GP, DP = net
    _, grads, _ = RN.phase_gmain(GP, DP, torch.from_numpy(z['z']), cfg, cfg)
    _check_grads(z, 'Gmain', 'G.', grads, 1e-4)


def test_phase_dmain_golden(net):
    z, meta, cfg, GP, DP = net
    _, grads = RN.phase_dmain(GP, DP, torch.from_numpy(z['z']), torch.from_numpy(z['real']), cfg, cfg)
    _check_grads(z, 'Dmain', 'D.', grads, 1e-4)


def test_phase_dreg_golden(net):
    z, meta, cfg, GP, DP = net
    _, grads, _ = RN.phase_dreg(DP, torch.from_numpy(z['real']), cfg, r1_gamma=meta['r1_gamma'], gain=meta['gains']['Dreg'])
    _check_grads(z, 'Dreg', 'D.', grads, 2e-4)


def test_phase_greg_golden(net):
    z, meta, cfg, GP, DP = net
    n = meta['n'] // 2
    _, grads, _ = RN.phase_greg(GP, torch.from_numpy(z['z'])[:n], torch.from_numpy(z['pl_noise']), cfg,
                                pl_weight=meta['pl_weight'], pl_decay=meta['pl_decay'], gain=meta['gains']['Greg'])
    _check_grads(z, 'Greg', 'G.', grads, 2e-4)
