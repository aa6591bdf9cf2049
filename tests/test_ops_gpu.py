"""GPU parity tests: the CUDA ops (through the C ABI) against the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star): 1e-4 relative for fp32 kernels, 1e-2 relative for fp16 / bf16 /
tensor-core paths, both measured against the fp32 oracle."""
import math

import pytest
import torch

from helpers import Golden, assert_close, rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu

F32_TOL = 1e-4
LOWP_TOL = 1e-2
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops():
    import sgb200
    from sgb200 import _lib
    assert _lib.lib().sgb_abi_version() == 1
    return sgb200.ops


def tol_for(dtype):
    return F32_TOL if dtype in (torch.float32, torch.float64) else LOWP_TOL


# ------------------------------------------------------------------------------------------------ bias_act
def test_bias_act_golden(ops):
    g = Golden('ops_bias_act.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x', DEV).requires_grad_(True)
        if c.get('nobias'):
            assert_close(ops.bias_act.bias_act(x, None, act=c['act']), g.t(cid, 'y'), F32_TOL, cid)
            continue
        b = g.t(cid, 'b', DEV).requires_grad_(True)
        kw = dict(dim=c['dim'], act=c['act'], alpha=c['alpha'], gain=c['gain'], clamp=c['clamp'])
        y = ops.bias_act.bias_act(x, b, **kw)
        assert_close(y, g.t(cid, 'y'), F32_TOL, f'{c} y')
        dy = g.t(cid, 'dy', DEV)
        dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
        assert_close(dx, g.t(cid, 'dx'), F32_TOL, f'{c} dx')
        assert_close(db, g.t(cid, 'db'), F32_TOL, f'{c} db')
        dyl = dy.clone().requires_grad_(True)
        dx2, = torch.autograd.grad(ops.bias_act.bias_act(x, b, **kw), [x], dyl, create_graph=True)
        g_dy, g_x = torch.autograd.grad((dx2 * g.t(cid, 'v', DEV)).sum(), [dyl, x], allow_unused=True)
        assert_close(g_dy, g.t(cid, 'g_dy'), F32_TOL, f'{c} g_dy')
        assert_close(g_x if g_x is not None else torch.zeros_like(x), g.t(cid, 'g_x'), 5e-4, f'{c} g_x')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16, torch.float64])
@pytest.mark.parametrize('shape,dim,cl', [
    ([4, 64, 32, 32], 1, False),     # vector path, plane bias
    ([4, 64, 32, 32], 1, True),      # channels_last: inner bias
    ([8, 512], 1, False),            # FC case
    ([3, 5, 7, 9], 1, False),        # odd sizes: generic bias + scalar tail
    ([2, 6, 5, 3], 1, True),
    ([2, 32, 4, 4], 1, False),
])
@pytest.mark.parametrize('act,gain,clamp', [('lrelu', None, None), ('lrelu', 1.0, 181.02), ('linear', None, 256.0),
                                            ('lrelu', math.sqrt(0.5), 0.4)])
def test_bias_act_shapes_dtypes(ops, dtype, shape, dim, cl, act, gain, clamp):
    torch.manual_seed(0)
    x32 = torch.randn(shape) * 2
    b32 = torch.randn(shape[dim])
    x = x32.to(DEV, dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    b = b32.to(DEV, dtype).requires_grad_(True)
    y = ops.bias_act.bias_act(x, b, dim=dim, act=act, gain=gain, clamp=clamp)
    assert y.dtype == dtype and y.shape == x.shape and y.stride() == x.stride()
    # oracle in fp32 (fp64 for fp64) on the values the kernel actually saw
    od = torch.float64 if dtype == torch.float64 else torch.float32
    xo = x.detach().cpu().to(od).requires_grad_(True)
    bo = b.detach().cpu().to(od).requires_grad_(True)
    yo = R.bias_act(xo, bo, dim=dim, act=act, gain=gain, clamp=clamp)
    tol = tol_for(dtype)
    assert_close(y, yo, tol, 'y')
    dy = torch.randn(shape).to(DEV, dtype)
    dx, db = torch.autograd.grad(y, [x, b], dy)
    dxo, dbo = torch.autograd.grad(yo, [xo, bo], dy.cpu().to(od))
    # The clamp derivative is decided from the STORED output (like the reference kernel, bias_act.cu:141), so in
    # 16-bit storage outputs within one ulp of the clamp value are indistinguishable from saturated ones.  That
    # boundary set has measure zero in exact arithmetic; exclude it from the max-norm comparison.
    safe = torch.ones_like(dxo, dtype=torch.bool)
    if clamp is not None and dtype in (torch.float16, torch.bfloat16):
        pre = R.bias_act(xo.detach(), bo.detach(), dim=dim, act=act, gain=gain, clamp=None).abs()
        safe = (pre - clamp).abs() > 4 * torch.finfo(dtype).eps * clamp
        assert safe.float().mean() > 0.95
    assert_close(dx.detach().cpu().to(od) * safe, dxo * safe, tol, 'dx')
    if bool(safe.all()):
        assert_close(db, dbo, 2 * tol if dtype not in (torch.float16, torch.bfloat16) else 3e-2, 'db')


def test_bias_act_unaligned_and_empty(ops):
    base = torch.randn(4 * 16 * 9 + 3, device=DEV)
    x = base[3:].reshape(4, 16, 3, 3)          # 12-byte offset: not 16B aligned
    b = torch.randn(16, device=DEV)
    assert_close(ops.bias_act.bias_act(x, b, act='lrelu'), R.bias_act(x.cpu(), b.cpu(), act='lrelu'), F32_TOL)
    e = torch.empty(0, 4, 3, 3, device=DEV)
    assert ops.bias_act.bias_act(e, torch.zeros(4, device=DEV), act='lrelu').shape == e.shape


def test_bias_act_errors(ops):
    with pytest.raises(RuntimeError):
        ops.bias_act.bias_act(torch.zeros(2, 3), torch.zeros(3))                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ops.bias_act.bias_act(torch.zeros(2, 3, device=DEV), torch.zeros(3, device=DEV, dtype=torch.float16))
    with pytest.raises(AssertionError):
        ops.bias_act.bias_act(torch.zeros(2, 3, device=DEV), torch.zeros(4, device=DEV))


def test_bias_act_gradcheck_fp64(ops):
    torch.manual_seed(1)
    for act in ['lrelu', 'linear', 'tanh', 'sigmoid', 'swish', 'softplus', 'elu', 'selu']:
        x = (torch.randn(2, 3, 4, 2, dtype=torch.float64, device=DEV) * 1.5).requires_grad_(True)
        b = torch.randn(3, dtype=torch.float64, device=DEV).requires_grad_(True)
        fn = lambda x, b: ops.bias_act.bias_act(x, b, act=act, gain=1.3)
        assert torch.autograd.gradcheck(fn, (x, b), eps=1e-6, atol=1e-6, nondet_tol=1e-9)
        assert torch.autograd.gradgradcheck(fn, (x, b), eps=1e-6, atol=1e-5, nondet_tol=1e-9)


# ----------------------------------------------------------------------------------------------- upfirdn2d
def test_upfirdn2d_golden(ops):
    g = Golden('ops_upfirdn2d.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x', DEV).requires_grad_(True)
        f = g.t(cid, 'f', DEV)
        if c['fn'] == 'upfirdn2d':
            y = ops.upfirdn2d.upfirdn2d(x, f, up=c['up'], down=c['down'], padding=c['padding'],
                                        flip_filter=c['flip_filter'], gain=c['gain'])
            assert_close(y, g.t(cid, 'y'), F32_TOL, f'{c} y')
            dx, = torch.autograd.grad(y, [x], g.t(cid, 'dy', DEV))
            assert_close(dx, g.t(cid, 'dx'), F32_TOL, f'{c} dx')
        else:
            y = getattr(ops.upfirdn2d, c['fn'])(x, f)
            assert_close(y, g.t(cid, 'y'), F32_TOL, f'{c} y')


def test_setup_filter(ops):
    assert_close(ops.upfirdn2d.setup_filter([1, 3, 3, 1]), R.setup_filter([1, 3, 3, 1]), 1e-7)
    assert ops.upfirdn2d.setup_filter([1] * 12).ndim == 1


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('case', [
    # (shape, up, down, padding, gain)  -- the four call-site forms at sizes that use the tiled kernels
    ([2, 8, 65, 65], 1, 1, [1, 1, 1, 1], 4),
    ([2, 3, 64, 64], 2, 1, [2, 1, 2, 1], 4),
    ([2, 8, 64, 64], 1, 1, [2, 2, 2, 2], 1),
    ([2, 8, 64, 64], 1, 2, [1, 1, 1, 1], 1),
    ([1, 4, 33, 47], 2, 1, [2, 1, 2, 1], 4),
    ([1, 4, 129, 70], 1, 2, [1, 1, 1, 1], 1),
    ([1, 4, 17, 16], 1, 1, [1, 1, 1, 1], 4),
])
def test_upfirdn2d_sizes(ops, dtype, cl, case):
    shape, up, down, pad, gain = case
    torch.manual_seed(2)
    f = R.setup_filter([1, 3, 3, 1])
    x = torch.randn(shape).to(DEV, dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = ops.upfirdn2d.upfirdn2d(x, f.to(DEV), up=up, down=down, padding=pad, gain=gain)
    xo = x.detach().cpu().float().requires_grad_(True)
    yo = R.upfirdn2d(xo, f, up=up, down=down, padding=pad, gain=gain)
    tol = tol_for(dtype)
    assert y.dtype == dtype
    assert y.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    assert_close(y, yo, tol, 'y')
    dy = torch.randn_like(yo)
    # first and second order: d<dx, v>/d(dy) must be upfirdn2d(v) again
    dyd = dy.to(DEV, dtype).requires_grad_(True)
    dx, = torch.autograd.grad(y, [x], dyd, create_graph=True)
    dyo = dy.clone().requires_grad_(True)
    dxo, = torch.autograd.grad(yo, [xo], dyo, create_graph=True)
    assert_close(dx, dxo, tol, 'dx')
    v = torch.randn_like(dxo)
    gg, = torch.autograd.grad((dx * v.to(DEV, dtype)).sum(), [dyd])
    ggo, = torch.autograd.grad((dxo * v).sum(), [dyo])
    assert_close(gg, ggo, tol, 'ddy')


def test_upfirdn2d_gradcheck_fp64(ops):
    torch.manual_seed(3)
    f = torch.randn(3, 4, device=DEV)
    x = torch.randn(1, 2, 6, 5, dtype=torch.float64, device=DEV, requires_grad=True)
    fn = lambda x: ops.upfirdn2d.upfirdn2d(x, f, up=[2, 1], down=[1, 2], padding=[1, 2, 0, 1], gain=1.7)
    assert torch.autograd.gradcheck(fn, (x,), eps=1e-6, atol=1e-6, nondet_tol=1e-9)
    assert torch.autograd.gradgradcheck(fn, (x,), eps=1e-6, atol=1e-6, nondet_tol=1e-9)


def test_upfirdn2d_errors(ops):
    x = torch.zeros(1, 1, 2, 2, device=DEV)
    with pytest.raises(RuntimeError):
        ops.upfirdn2d.upfirdn2d(x, torch.ones(4, 4, device=DEV))           # output smaller than 1x1
    with pytest.raises(RuntimeError):
        ops.upfirdn2d.upfirdn2d(x, torch.ones(1, 1, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        ops.upfirdn2d.upfirdn2d(x.cpu(), torch.ones(1, 1))


# ----------------------------------------------------------------------------------------- conv2d_resample
def test_conv2d_resample_golden(ops):
    g = Golden('ops_conv.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x', DEV).requires_grad_(True)
        w = g.t(cid, 'w', DEV).requires_grad_(True)
        y = ops.conv2d_resample.conv2d_resample(x, w, f=g.t(cid, 'f', DEV), up=c['up'], down=c['down'], padding=c['padding'],
                                                groups=c['groups'], flip_weight=c['flip_weight'])
        assert_close(y, g.t(cid, 'y'), F32_TOL, f'{c} y')
        dx, dw = torch.autograd.grad(y, [x, w], g.t(cid, 'dy', DEV))
        assert_close(dx, g.t(cid, 'dx'), F32_TOL, f'{c} dx')
        assert_close(dw, g.t(cid, 'dw'), F32_TOL, f'{c} dw')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, k, up, down, groups)
    (2, 32, 48, 16, 3, 1, 1, 1),
    (2, 32, 32, 16, 3, 2, 1, 1),
    (2, 32, 48, 16, 3, 1, 2, 1),
    (2, 64, 3, 16, 1, 1, 1, 1),
    (2, 3, 32, 16, 1, 1, 1, 1),
    (2, 32, 32, 16, 1, 1, 2, 1),
    (3, 8, 12, 9, 3, 1, 1, 1),
    (2, 16, 24, 8, 3, 2, 1, 2),
    (1, 130, 70, 7, 3, 1, 1, 1),
])
def test_conv2d_resample_sizes(ops, dtype, cl, case):
    n, ci, co, h, k, up, down, groups = case
    torch.manual_seed(4)
    f = R.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, ci, h, h).to(DEV, dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    w = (torch.randn(co, ci // groups, k, k) / math.sqrt(ci * k * k)).to(DEV, dtype).requires_grad_(True)
    fw = (up == 1)
    y = ops.conv2d_resample.conv2d_resample(x, w, f=f.to(DEV), up=up, down=down, padding=k // 2, groups=groups, flip_weight=fw)
    xo = x.detach().cpu().float().requires_grad_(True)
    wo = w.detach().cpu().float().requires_grad_(True)
    yo = R.conv2d_resample(xo, wo, f=f, up=up, down=down, padding=k // 2, groups=groups, flip_weight=fw)
    tol = tol_for(dtype)
    assert_close(y, yo, tol, 'y')
    dy = torch.randn_like(yo)
    dx, dw = torch.autograd.grad(y, [x, w], dy.to(DEV, dtype))
    dxo, dwo = torch.autograd.grad(yo, [xo, wo], dy)
    assert_close(dx, dxo, tol, 'dx')
    assert_close(dw, dwo, tol, 'dw')


def test_conv_no_weight_gradients_and_r1_double_backward(ops):
    """R1-style: grad of |d logits / d x|^2 wrt the weights, first-order pass under no_weight_gradients()."""
    torch.manual_seed(5)
    f = R.setup_filter([1, 3, 3, 1])
    x = torch.randn(2, 4, 8, 8, device=DEV, requires_grad=True)
    w1 = (torch.randn(6, 4, 3, 3, device=DEV) / 6).requires_grad_(True)
    w2 = (torch.randn(5, 6, 3, 3, device=DEV) / 7).requires_grad_(True)
    b = torch.randn(6, device=DEV).requires_grad_(True)

    def net(cr, ba, x, w1, w2, b, f):
        h = cr(x, w1, f=f, padding=1)
        h = ba(h, b, act='lrelu')
        h = cr(h, w2, f=f, down=2, padding=1)
        return h.square().sum()

    out = net(ops.conv2d_resample.conv2d_resample, ops.bias_act.bias_act, x, w1, w2, b, f.to(DEV))
    with ops.conv2d_gradfix.no_weight_gradients():
        gx, = torch.autograd.grad(out, [x], create_graph=True)
    pen = gx.square().sum()
    gw1, gw2, gb = torch.autograd.grad(pen, [w1, w2, b])
    xo, w1o, w2o, bo = [t.detach().cpu().requires_grad_(True) for t in (x, w1, w2, b)]
    outo = net(R.conv2d_resample, R.bias_act, xo, w1o, w2o, bo, f)
    gxo, = torch.autograd.grad(outo, [xo], create_graph=True)
    gw1o, gw2o, gbo = torch.autograd.grad(gxo.square().sum(), [w1o, w2o, bo])
    assert_close(gx, gxo, F32_TOL, 'gx')
    assert_close(gw1, gw1o, 2e-4, 'gw1')
    assert_close(gw2, gw2o, 2e-4, 'gw2')
    assert_close(gb, gbo, 2e-4, 'gb')
    # the flag really suppresses the weight gradient of the first-order pass
    out2 = net(ops.conv2d_resample.conv2d_resample, ops.bias_act.bias_act, x, w1, w2, b, f.to(DEV))
    with ops.conv2d_gradfix.no_weight_gradients():
        res = torch.autograd.grad(out2, [x, w1], allow_unused=True)
    assert res[1] is None


def test_conv_gradcheck_fp64(ops):
    torch.manual_seed(6)
    f = R.setup_filter([1, 3, 3, 1]).to(DEV)
    for up, down, k in [(1, 1, 3), (2, 1, 3), (1, 2, 3), (1, 1, 1)]:
        x = torch.randn(1, 2, 5, 5, dtype=torch.float64, device=DEV, requires_grad=True)
        w = torch.randn(3, 2, k, k, dtype=torch.float64, device=DEV, requires_grad=True)
        fn = lambda x, w: ops.conv2d_resample.conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
        assert torch.autograd.gradcheck(fn, (x, w), eps=1e-6, atol=1e-6, nondet_tol=1e-7)
        assert torch.autograd.gradgradcheck(fn, (x, w), eps=1e-6, atol=1e-5, nondet_tol=1e-7)


# --------------------------------------------------------------------------------------- modulated_conv2d
def test_modulated_conv2d_golden(ops):
    from sgb200.modconv import modulated_conv2d
    g = Golden('ops_modconv.npz')
    for c in g.cases():
        cid = c['id']
        x = g.t(cid, 'x', DEV).requires_grad_(True)
        w = g.t(cid, 'w', DEV).requires_grad_(True)
        s = g.t(cid, 's', DEV).requires_grad_(True)
        y = modulated_conv2d(x, w, s, noise=g.t(cid, 'noise', DEV), up=c['up'], padding=c['padding'],
                             resample_filter=g.t(cid, 'f', DEV), demodulate=c['demodulate'],
                             flip_weight=c['flip_weight'], fused_modconv=c['fused_modconv'])
        assert_close(y, g.t(cid, 'y'), F32_TOL, f'{c} y')
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], g.t(cid, 'dy', DEV), create_graph=True)
        assert_close(dx, g.t(cid, 'dx'), F32_TOL, f'{c} dx')
        assert_close(dw, g.t(cid, 'dw'), F32_TOL, f'{c} dw')
        assert_close(ds, g.t(cid, 'ds'), F32_TOL, f'{c} ds')
        g2w, g2s = torch.autograd.grad(ds.square().sum(), [w, s], allow_unused=True)
        for got, key in ((g2w, 'g2w'), (g2s, 'g2s')):
            want = g.t(cid, key)
            if want is None:
                assert got is None or float(got.abs().max()) == 0.0, f'{c} {key} presence'
            else:
                assert_close(got, want, 5e-4, f'{c} {key}')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('up,k,demod', [(1, 3, True), (2, 3, True), (1, 1, False)])
def test_modulated_conv2d_dtypes(ops, dtype, cl, up, k, demod):
    from sgb200.modconv import modulated_conv2d
    torch.manual_seed(7)
    n, ci, co, h = 3, 32, 32 if k == 3 else 3, 16
    f = R.setup_filter([1, 3, 3, 1])
    x = torch.randn(n, ci, h, h).to(DEV, dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    w = torch.randn(co, ci, k, k, device=DEV, requires_grad=True)
    s = (torch.randn(n, ci, device=DEV) + 1).requires_grad_(True)
    if not demod:
        s = (s.detach() / math.sqrt(ci)).requires_grad_(True)
    noise = torch.randn(n, 1, h * up, h * up, device=DEV) * 0.2 if demod else None
    y = modulated_conv2d(x, w, s, noise=noise, up=up, padding=k // 2, resample_filter=f.to(DEV), demodulate=demod,
                         flip_weight=(up == 1), fused_modconv=False)
    xo = x.detach().cpu().float().requires_grad_(True)
    wo = w.detach().cpu().requires_grad_(True)
    so = s.detach().cpu().requires_grad_(True)
    yo = R.modulated_conv2d(xo, wo, so, noise=None if noise is None else noise.cpu(), up=up, padding=k // 2,
                            resample_filter=f, demodulate=demod, flip_weight=(up == 1), fused_modconv=False)
    tol = tol_for(dtype)
    assert y.dtype == dtype
    assert_close(y, yo, tol, 'y')
    dy = torch.randn_like(yo)
    dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy.to(DEV, dtype))
    dxo, dwo, dso = torch.autograd.grad(yo, [xo, wo, so], dy)
    assert_close(dx, dxo, tol, 'dx')
    assert_close(dw, dwo, tol, 'dw')
    assert_close(ds, dso, tol, 'ds')


def test_fma_and_reductions(ops):
    torch.manual_seed(8)
    for cl in (False, True):
        a = torch.randn(3, 10, 6, 7, device=DEV)
        if cl:
            a = a.contiguous(memory_format=torch.channels_last)
        a.requires_grad_(True)
        b = torch.randn(3, 10, 1, 1, device=DEV, requires_grad=True)
        c = torch.randn(3, 1, 6, 7, device=DEV, requires_grad=True)
        y = ops.fma.fma(a, b, c)
        ao, bo, co = [t.detach().cpu().requires_grad_(True) for t in (a, b, c)]
        yo = R.fma(ao, bo, co)
        assert_close(y, yo, F32_TOL, 'fma')
        dy = torch.randn_like(yo)
        ga, gb, gc = torch.autograd.grad(y, [a, b, c], dy.to(DEV))
        gao, gbo, gco = torch.autograd.grad(yo, [ao, bo, co], dy)
        assert_close(ga, gao, F32_TOL, 'da')
        assert_close(gb, gbo, F32_TOL, 'db')
        assert_close(gc, gco, F32_TOL, 'dc')
    a = torch.randn(2, 3, 4, 5, dtype=torch.float64, device=DEV, requires_grad=True)
    b = torch.randn(2, 3, 1, 1, dtype=torch.float64, device=DEV, requires_grad=True)
    c = torch.randn(2, 1, 4, 5, dtype=torch.float64, device=DEV, requires_grad=True)
    assert torch.autograd.gradcheck(ops.fma.fma, (a, b, c), eps=1e-6, atol=1e-6, nondet_tol=1e-9)
    assert torch.autograd.gradgradcheck(ops.fma.fma, (a, b, c), eps=1e-6, atol=1e-6, nondet_tol=1e-9)
