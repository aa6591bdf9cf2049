"""GPU: conv2d with the fused epilogue (demodulation scale, noise, bias, lrelu|linear, gain, clamp in the tcgen05
kernel; one-pass backward) against the un-fused composition of the same ops and against a plain PyTorch fp32
restatement of generators.py:80-87 + bias_act.py:93-123 on the CPU.  Tolerance 1e-2 (tensor-core class)."""
import math

import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-2


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _l2_close(a, b, tol, what):
    """relative L2 error: robust against the isolated elements whose pre-activation lies within rounding distance of an
    lrelu / clamp breakpoint and therefore takes the other branch than the fp32 CPU evaluation (one such element moves a
    max-norm comparison by several percent when a third of the outputs are saturated)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    e = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert e <= tol, f'{what}: relative L2 error {e:.3e} > {tol:.1e}'


def _cpu_ref(x, w, b, s, d, nz, stride, pad, act, gain, clamp):
    y = torch.nn.functional.conv2d(x * s[:, :, None, None] if s is not None else x, w, stride=stride, padding=pad)
    if d is not None:
        y = y * d[:, :, None, None]
    if nz is not None:
        y = y + nz
    if b is not None:
        y = y + b[None, :, None, None]
    if act == 'lrelu':
        y = torch.nn.functional.leaky_relu(y, 0.2)
    y = y * gain
    if clamp is not None:
        y = y.clamp(-clamp, clamp)
    return y


def _inputs(dtype, n, ci, co, h, k, with_mod, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, ci, h, h, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
    b = torch.randn(co, generator=g) * 0.5
    s = (torch.randn(n, ci, generator=g) * 0.3 + 1) if with_mod else None
    d = (torch.rand(n, co, generator=g) + 0.5) if with_mod else None
    nz = (torch.randn(n, 1, (h + 2 * (k // 2) - k) + 1, (h + 2 * (k // 2) - k) + 1, generator=g) * 0.3) if with_mod else None
    return x, w, b, s, d, nz


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
@pytest.mark.parametrize('cfg', [
    # (n, ci, co, h, k, stride, with_mod, act, gain, clamp)
    (3, 64, 32, 16, 3, 1, True, 'lrelu', math.sqrt(2), None),
    (3, 64, 32, 16, 3, 1, True, 'lrelu', math.sqrt(2), 0.7),        # clamp saturates many outputs
    (2, 32, 64, 16, 1, 1, True, 'linear', 1.0, 0.9),                # toRGB-like (1x1, no demod would pass d = None)
    (2, 64, 64, 17, 3, 2, False, 'lrelu', 1.0, 181.02 / 200),       # D conv1 (stride 2 after the FIR), gain folded
    (2, 64, 128, 16, 1, 1, False, 'linear', math.sqrt(0.5), None),  # D skip
    (2, 64, 8, 8, 3, 1, True, 'lrelu', math.sqrt(2), 256.0),        # small co (one 16-byte vector in fp16)
])
def test_fused_conv_matches_unfused_and_cpu(dtype, cfg):
    from sgb200.ops import fused_conv, conv2d_gradfix as cg
    n, ci, co, h, k, stride, with_mod, act, gain, clamp = cfg
    torch.backends.cudnn.allow_tf32 = True
    pad = k // 2 if stride == 1 else 0
    x, w, b, s, d, nz = _inputs(dtype, n, ci, co, h, k, with_mod and stride == 1)
    if nz is not None and stride != 1:
        nz = None
    leaves_cpu = [t.clone().requires_grad_(True) if t is not None else None for t in (x, w, b, s, d, nz)]
    yo = _cpu_ref(*leaves_cpu, stride, pad, act, gain, clamp)
    dy = torch.randn(yo.shape, generator=torch.Generator().manual_seed(9))
    go = torch.autograd.grad(yo, [t for t in leaves_cpu if t is not None], dy)

    def run(fused):
        fused_conv.enabled = fused
        try:
            lv = []
            for i, t in enumerate((x, w, b, s, d, nz)):
                if t is None:
                    lv.append(None)
                    continue
                tt = t.to(DEV, dtype if i in (0, 1, 2) else torch.float32)
                if i == 0:
                    tt = _cl(tt)
                lv.append(tt.requires_grad_(True))
            y = fused_conv.conv2d_bias_act(lv[0], lv[1], lv[2], stride=stride, padding=pad, styles=lv[3], dcoefs=lv[4],
                                           noise=lv[5], act=act, gain=gain, clamp=clamp)
            gs = torch.autograd.grad(y, [t for t in lv if t is not None], _cl(dy.to(DEV, dtype)))
            return y, gs
        finally:
            fused_conv.enabled = False

    yf, gf = run(True)
    yu, gu = run(False)
    fused_conv.enabled = False
    assert yf.is_contiguous(memory_format=torch.channels_last)
    assert_close(yf, yo, TOL, f'{cfg} {dtype} y vs cpu')
    assert_close(yf, yu.float().cpu(), 4e-3, f'{cfg} {dtype} y fused vs unfused')
    names = [nm for nm, t in zip('x w b styles dcoefs noise'.split(), (x, w, b, s, d, nz)) if t is not None]
    for nm, a, u, o in zip(names, gf, gu, go):
        assert a.shape == o.shape
        # a clamp that saturates a third of the outputs puts many elements within rounding distance of the breakpoint
        _l2_close(a, o, 5e-2 if (clamp is not None and clamp < 1) else TOL, f'{cfg} {dtype} d{nm} vs cpu')
        # fp16: the un-fused path rounds the pre-activation to fp16 before lrelu, so a handful of near-zero elements pick the
        # other slope than the fused / fp32 evaluation; each flips one dz entry by 0.8 * |dy| * gain
        if dtype == torch.float16:
            _l2_close(a, u, 5e-2 if (clamp is not None and clamp < 1) else 2e-2, f'{cfg} {dtype} d{nm} fused vs unfused')
        else:       # same TF32 convolution, fp32 epilogue on both routes: the two must agree closely
            assert_close(a, u.float().cpu(), 2e-3, f'{cfg} {dtype} d{nm} fused vs unfused')


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
def test_fused_conv_second_order(dtype):
    """create_graph=True routes the backward through the differentiable composition: R1-style (penalty on dL/dx
    differentiated wrt the weights) and path-length-style (dL/dstyles differentiated wrt w and styles) terms."""
    from sgb200.ops import fused_conv, conv2d_gradfix as cg
    torch.backends.cudnn.allow_tf32 = True
    x, w, b, s, d, nz = _inputs(dtype, 2, 32, 32, 8, 3, True, seed=3)

    def second_order(mk, conv_fn):
        lx, lw, lb, ls, ld, ln = mk(x), mk(w), mk(b), mk(s), mk(d), mk(nz)
        y = conv_fn(lx, lw, lb, ls, ld, ln)
        gx, gs = torch.autograd.grad(y.float().square().sum() * 0.01, [lx, ls], create_graph=True)
        pen = gx.float().square().sum() + gs.float().square().sum()
        return torch.autograd.grad(pen, [lw, ls, lb])

    ref = second_order(lambda t: t.clone().double().requires_grad_(True),
                       lambda lx, lw, lb, ls, ld, ln: _cpu_ref(lx, lw, lb, ls, ld, ln, 1, 1, 'lrelu', math.sqrt(2), None))
    idx = {id(t): i for i, t in enumerate((x, w, b, s, d, nz))}

    def mk_dev(t):
        i = idx[id(t)]
        tt = t.to(DEV, dtype if i in (0, 1, 2) else torch.float32)
        return (_cl(tt) if i == 0 else tt).requires_grad_(True)

    fused_conv.enabled = True
    try:
        got = second_order(mk_dev, lambda lx, lw, lb, ls, ld, ln: fused_conv.conv2d_bias_act(
            lx, lw, lb, stride=1, padding=1, styles=ls, dcoefs=ld, noise=ln, act='lrelu', gain=math.sqrt(2)))
    finally:
        fused_conv.enabled = False
    for nm, a, o in zip(('w', 'styles', 'b'), got, ref):
        assert_close(a, o.float(), 2e-2 if dtype == torch.float16 else TOL, f'second order d{nm} {dtype}')


def test_fused_conv_honours_no_weight_gradients():
    from sgb200.ops import fused_conv, conv2d_gradfix as cg
    torch.backends.cudnn.allow_tf32 = True
    x, w, b, s, d, nz = _inputs(torch.float16, 2, 32, 32, 8, 3, False)
    lx = _cl(x.to(DEV, torch.float16)).requires_grad_(True)
    lw = w.to(DEV, torch.float16).requires_grad_(True)
    fused_conv.enabled = True
    try:
        y = fused_conv.conv2d_bias_act(lx, lw, b.to(DEV, torch.float16), padding=1, act='lrelu')
        with cg.no_weight_gradients():
            gx, gw = torch.autograd.grad(y.sum(), [lx, lw], allow_unused=True)
    finally:
        fused_conv.enabled = False
    assert gw is None and gx is not None


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
@pytest.mark.parametrize('clamp', [None, 0.8])
def test_scale_bias_act_tail_matches_unfused(dtype, clamp):
    """One-pass demodulation + noise + bias_act (and its one-pass backward) vs the fma -> bias_act composition."""
    from sgb200.ops import fused_conv
    g = torch.Generator().manual_seed(4)
    n, c, h = 3, 32, 12
    x = torch.randn(n, c, h, h, generator=g)
    b = torch.randn(c, generator=g) * 0.3
    d = torch.rand(n, c, generator=g) + 0.5
    nz = torch.randn(n, 1, h, h, generator=g) * 0.2
    dy = torch.randn(n, c, h, h, generator=g)

    def run(on):
        fused_conv.tail_enabled = on
        try:
            lx = _cl(x.to(DEV, dtype)).requires_grad_(True)
            lb = b.to(DEV, dtype).requires_grad_(True)
            ld = d.to(DEV).requires_grad_(True)
            ln = nz.to(DEV).requires_grad_(True)
            y = fused_conv.scale_bias_act(lx, lb, ld, ln, act='lrelu', gain=math.sqrt(2), clamp=clamp)
            gs = torch.autograd.grad(y, [lx, lb, ld, ln], _cl(dy.to(DEV, dtype)))
            return y, gs
        finally:
            fused_conv.tail_enabled = True

    yf, gf = run(True)
    yu, gu = run(False)
    # fp16: the un-fused route rounds x * dcoefs + noise to fp16 before bias_act, so elements within rounding distance of
    # the lrelu / clamp breakpoints take the other branch (see _l2_close); a saturating clamp puts many elements there
    tol = (6e-2 if clamp is not None else 2e-2) if dtype == torch.float16 else 1e-4
    assert_close(yf, yu.float().cpu(), 2e-3 if dtype == torch.float16 else 1e-5, f'y {dtype}')
    for nm, a, u in zip('x b dcoefs noise'.split(), gf, gu):
        _l2_close(a, u, tol, f'd{nm} {dtype} clamp={clamp}')
    # second order through the differentiable route
    lx = _cl(x.to(DEV, dtype)).requires_grad_(True)
    ld = d.to(DEV).requires_grad_(True)
    y = fused_conv.scale_bias_act(lx, b.to(DEV, dtype), ld, nz.to(DEV), act='lrelu', gain=math.sqrt(2), clamp=clamp)
    gx, = torch.autograd.grad(y.float().square().sum(), [lx], create_graph=True)
    gd, = torch.autograd.grad(gx.float().square().sum(), [ld])
    assert torch.isfinite(gd).all() and gd.abs().max() > 0
