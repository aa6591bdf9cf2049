"""GPU: the tcgen05 implicit-GEMM convolution (channels_last, f16 / bf16 / TF32) against the CPU oracle, and
against the SIMT kernel on the same inputs.  Tolerance 1e-2 relative (tensor-core class, BASELINE.json)."""
import math

import pytest
import torch

from helpers import assert_close, rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-2


def _desc_uses_tc(x, w, transposed=False, stride=1, pad=0):
    from sgb200.ops import conv2d_gradfix as cg
    from sgb200 import _lib
    n, ci, h, wd = x.shape
    if not transposed:
        co, kh, kw = w.shape[0], w.shape[2], w.shape[3]
        oh, ow = (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
    else:
        co, kh, kw = w.shape[1], w.shape[2], w.shape[3]
        oh, ow = (h - 1) * stride - 2 * pad + kh, (wd - 1) * stride - 2 * pad + kw
    y = torch.empty([n, co, oh, ow], dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    d = cg._make_desc(x, y, transposed, ci, co, kh, kw, stride, (pad, pad), 1, False)
    ws = cg._attach_workspace(d, x.device)
    return bool(_lib.lib().sgb_conv2d_uses_tensor_cores(d))


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def test_probe_1x1_gemm_fp16():
    """Smallest case: a 1x1 convolution is a plain [pixels x ci] x [ci x co] GEMM -> checks the UMMA descriptors."""
    from sgb200.ops import conv2d_gradfix as cg
    torch.manual_seed(0)
    x = _cl(torch.randn(1, 64, 16, 16).to(DEV, torch.float16))
    w = (torch.randn(64, 64, 1, 1) / 8).to(DEV, torch.float16)
    assert _desc_uses_tc(x, w)
    y = cg.conv2d(x, w)
    yo = torch.nn.functional.conv2d(x.cpu().float(), w.cpu().float())
    assert_close(y, yo, TOL, '1x1 fp16')


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, W, k, stride, pad, transposed)
    (1, 64, 64, 16, 16, 1, 1, 0, False),
    (2, 64, 64, 16, 16, 3, 1, 1, False),
    (2, 32, 48, 12, 20, 3, 1, 1, False),      # M tail, co not a power of two
    (3, 72, 24, 9, 9, 3, 1, 1, False),        # K tail (72 = 64 + 8), several samples per tile
    (2, 64, 128, 17, 17, 3, 2, 0, False),     # stride 2 (D down path, after the FIR)
    (2, 128, 64, 8, 8, 3, 1, 1, True),        # transposed stride 1 (dgrad of a 3x3 conv)
    (2, 64, 32, 8, 8, 3, 2, 0, True),         # transposed stride 2 (G up path)
    (1, 512, 512, 4, 4, 3, 1, 1, False),      # N tiles > 1, tiny image
    (2, 64, 3, 16, 16, 1, 1, 0, False),       # toRGB: co = 3
    (1, 256, 272, 6, 6, 3, 1, 1, False),      # co tail over two N tiles
])
def test_umma_conv_vs_oracle(dtype, case):
    from sgb200.ops import conv2d_gradfix as cg
    n, ci, co, h, wd, k, stride, pad, tr = case
    torch.manual_seed(1)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(n, ci, h, wd).to(DEV, dtype))
    wshape = (ci, co, k, k) if tr else (co, ci, k, k)
    w = (torch.randn(wshape) / math.sqrt(ci * k * k)).to(DEV, dtype)
    assert _desc_uses_tc(x, w, tr, stride, pad)
    op = cg.conv_transpose2d if tr else cg.conv2d
    y = op(x, w, stride=stride, padding=pad)
    fo = torch.nn.functional.conv_transpose2d if tr else torch.nn.functional.conv2d
    yo = fo(x.cpu().float(), w.cpu().float(), stride=stride, padding=pad)
    assert y.shape == yo.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert_close(y, yo, TOL, f'{case} {dtype}')
    # same inputs through the SIMT kernel: the two CUDA paths must agree as well
    cg.use_tensor_cores = False
    try:
        ys = op(x, w, stride=stride, padding=pad)
    finally:
        cg.use_tensor_cores = True
    assert_close(y, ys.float().cpu(), TOL, f'{case} {dtype} vs simt')


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
def test_umma_in_scale_and_flip(dtype):
    from sgb200.ops import conv2d_gradfix as cg
    torch.manual_seed(2)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(3, 64, 8, 8).to(DEV, dtype))
    w = (torch.randn(32, 64, 3, 3) / 24).to(DEV, dtype)
    s = (torch.randn(3, 64) + 1).to(DEV)
    y = cg.conv2d(x, w, padding=1, flip_weight=True, in_scale=s)
    yo = torch.nn.functional.conv2d(x.cpu().float() * s.cpu()[:, :, None, None], w.cpu().float().flip([2, 3]), padding=1)
    assert_close(y, yo, TOL, 'in_scale + flip')


def test_umma_gradients_fp16():
    """dgrad runs through the transposed tensor-core kernel; compare all gradients with the oracle."""
    from sgb200.ops import conv2d_resample
    torch.manual_seed(3)
    f = R.setup_filter([1, 3, 3, 1])
    for up, down in [(1, 1), (2, 1), (1, 2)]:
        x = _cl(torch.randn(2, 64, 16, 16).to(DEV, torch.float16)).requires_grad_(True)
        w = (torch.randn(64, 64, 3, 3) / 24).to(DEV, torch.float16).requires_grad_(True)
        y = conv2d_resample.conv2d_resample(x, w, f=f.to(DEV), up=up, down=down, padding=1, flip_weight=(up == 1))
        xo = x.detach().cpu().float().requires_grad_(True)
        wo = w.detach().cpu().float().requires_grad_(True)
        yo = R.conv2d_resample(xo, wo, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))
        assert_close(y, yo, TOL, f'y up{up} down{down}')
        dy = torch.randn_like(yo)
        dx, dw = torch.autograd.grad(y, [x, w], dy.to(DEV, torch.float16))
        dxo, dwo = torch.autograd.grad(yo, [xo, wo], dy)
        assert_close(dx, dxo, TOL, f'dx up{up} down{down}')
        assert_close(dw, dwo, TOL, f'dw up{up} down{down}')


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, k, stride, pad)
    (2, 64, 64, 16, 3, 1, 1),
    (3, 32, 48, 9, 3, 1, 1),          # pixel tail, co/ci tails inside a tile
    (2, 64, 128, 17, 3, 2, 0),        # stride 2
    (1, 512, 512, 4, 3, 1, 1),        # several o / c tiles, 16 pixels only
    (4, 128, 8, 16, 1, 1, 0),         # 1x1, tiny co
    (2, 72, 264, 8, 3, 1, 1),         # co > 256 with a tail
    (2, 64, 64, 40, 3, 1, 1),         # several row tiles per image, last one partial
    (3, 128, 136, 20, 3, 1, 1),       # W = 20: partial column tile; two o tiles (136 = 128 + 8)
    (2, 40, 64, 33, 3, 2, 0),         # stride 2, odd input (G up / D down weight gradients), ci tail
    (2, 160, 32, 12, 3, 1, 1),        # two c tiles with a tail (160 = 128 + 32)
    (2, 64, 128, 16, 1, 1, 0),        # 1x1 (D skip)
    (70, 32, 32, 8, 3, 1, 1),         # more tiles than CTAs want: several tiles per CTA, pipeline wrap-around
    (2, 64, 32, 24, 3, 1, 1),         # one dy block (co = 32): filter rows stacked in M, 24 rows = 1.5 tiles of 16
    (1, 48, 64, 70, 3, 1, 0),         # no padding (out = in - 2), partial column tile, ci tail; many row tiles
    (5, 64, 64, 128, 3, 1, 1),        # a real layer size: 128 x 128, every CTA walks many tiles
])
def test_umma_wgrad_vs_oracle(dtype, case):
    from sgb200.ops import conv2d_gradfix as cg
    n, ci, co, h, k, stride, pad = case
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(n, ci, h, h).to(DEV, dtype))
    w = (torch.randn(co, ci, k, k) / math.sqrt(ci * k * k)).to(DEV, dtype).requires_grad_(True)
    s = (torch.randn(n, ci) + 1).to(DEV)
    for scale in (None, s):
        from sgb200 import _lib
        y = cg.conv2d(x, w, stride=stride, padding=pad, in_scale=scale)
        dy = _cl(torch.randn(y.shape).to(DEV, dtype))          # gradients arrive channels_last in the networks
        _lib.profile_start()
        dw, = torch.autograd.grad(y, [w], dy)
        torch.cuda.synchronize()
        assert set(_lib.profile_stop().summary()) == {'conv_wgrad_tc'}
        xo = x.cpu().float() * (1 if scale is None else scale.cpu()[:, :, None, None])
        wo = w.detach().cpu().float().requires_grad_(True)
        yo = torch.nn.functional.conv2d(xo, wo, stride=stride, padding=pad)
        dwo, = torch.autograd.grad(yo, [wo], dy.cpu().float())
        assert_close(dw, dwo, TOL, f'{case} {dtype} scale={scale is not None}')
        # the per-tap kernel (first tensor-core wgrad) must agree with the halo-tile kernel
        cg.use_halo_kernel = False
        try:
            dw1, = torch.autograd.grad(cg.conv2d(x, w, stride=stride, padding=pad, in_scale=scale), [w], dy)
        finally:
            cg.use_halo_kernel = True
        assert_close(dw, dw1.float().cpu(), 1e-2 if dtype == torch.bfloat16 else 3e-3, f'{case} {dtype} halo vs per-tap wgrad')


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, W, k, pad, transposed, stride)   -- geometries of the halo-tile kernel
    (1, 64, 64, 16, 16, 3, 1, False, 1),
    (3, 64, 64, 8, 8, 3, 1, False, 1),           # tiles straddle images (virtual rows)
    (2, 32, 48, 24, 16, 3, 1, False, 1),         # H not a multiple of 16, co tail
    (2, 136, 24, 16, 8, 3, 1, False, 1),         # K tail (136 = 2*64 + 8)
    (2, 128, 64, 16, 16, 3, 1, True, 1),         # dgrad form
    (2, 64, 64, 16, 16, 3, 0, True, 1),          # transposed, pad 0 (output grows by 2)
    (2, 64, 64, 18, 18, 3, 0, False, 1),         # valid conv: out 16x16
    (2, 64, 3, 16, 16, 1, 0, False, 1),          # toRGB
    (1, 512, 512, 8, 8, 3, 1, False, 1),         # two N tiles
    (1, 64, 272, 16, 16, 1, 0, False, 1),
    (5, 64, 64, 32, 32, 3, 1, False, 1),
    (4, 512, 512, 4, 4, 3, 1, False, 1),         # W = 4: partial column tile
    (2, 64, 64, 12, 20, 3, 1, False, 1),         # W = 20: partial last column tile
    (2, 64, 128, 33, 33, 3, 0, False, 2),        # stride 2: D down path after the FIR (H+1 -> H/2)
    (3, 40, 72, 17, 25, 3, 0, False, 2),         # stride 2, odd sizes, K tail (40 = 2*16 + 8), co tail
    (2, 64, 64, 16, 16, 3, 1, False, 2),         # stride 2 with padding
    (1, 24, 512, 9, 9, 3, 0, False, 2),          # stride 2, two N tiles
    (2, 128, 64, 16, 16, 3, 0, True, 2),         # transposed stride 2: G up path (H -> 2H+1)
    (3, 72, 40, 7, 11, 3, 0, True, 2),           # transposed stride 2, odd sizes, tails
    (2, 64, 64, 16, 16, 3, 1, True, 2),          # transposed stride 2 with padding (H -> 2H-1)
    (1, 64, 512, 8, 8, 3, 0, True, 2),           # transposed stride 2, co = 512 (four N tiles of 128)
    (5, 32, 16, 40, 24, 3, 0, True, 2),          # several row tiles per image
])
def test_halo_conv_vs_oracle_and_v1(dtype, case):
    from sgb200.ops import conv2d_gradfix as cg
    n, ci, co, h, wd, k, pad, tr, stride = case
    torch.manual_seed(11)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(n, ci, h, wd).to(DEV, dtype))
    wshape = (ci, co, k, k) if tr else (co, ci, k, k)
    w = (torch.randn(wshape) / math.sqrt(ci * k * k)).to(DEV, dtype)
    s = (torch.randn(n, ci) + 1).to(DEV)
    op = cg.conv_transpose2d if tr else cg.conv2d
    fo = torch.nn.functional.conv_transpose2d if tr else torch.nn.functional.conv2d
    for scale in (None, s):
        xo = x.cpu().float() * (1 if scale is None else scale.cpu()[:, :, None, None])
        yo = fo(xo, w.cpu().float(), stride=stride, padding=pad)
        for gt in (0, 1, 2, 4):          # tiles per super-tile: automatic, then forced (falls back to what TMEM allows)
            cg.halo_gt = gt
            try:
                y = op(x, w, stride=stride, padding=pad, in_scale=scale)
            finally:
                cg.halo_gt = 0
            assert y.shape == yo.shape
            assert_close(y, yo, TOL, f'{case} {dtype} halo gt={gt} scale={scale is not None}')
        cg.use_halo_kernel = False
        try:
            y1 = op(x, w, stride=stride, padding=pad, in_scale=scale)
        finally:
            cg.use_halo_kernel = True
        assert_close(y, y1.float().cpu(), {torch.float32: 1e-3, torch.float16: 2e-3, torch.bfloat16: 8e-3}[dtype],
                     f'{case} {dtype} halo vs per-tap kernel')


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, k, pad, nchw_input)
    (4, 3, 64, 32, 1, 0, True),        # fromRGB: NCHW image in, channels_last features out
    (4, 3, 64, 32, 1, 0, False),
    (4, 64, 3, 32, 1, 0, False),       # toRGB
    (8, 513, 512, 4, 3, 1, False),     # D epilogue conv after the minibatch-stddev channel
    (2, 5, 7, 9, 3, 1, False),         # both channel counts odd
])
def test_channel_padding_makes_small_channel_convs_tensor_core(dtype, case):
    """Convolutions whose channel count is not a multiple of one 16-byte vector are zero-padded in the op so that
    forward, data gradient and weight gradient all run on the tensor-core kernels; results vs the CPU oracle."""
    from sgb200.ops import conv2d_gradfix as cg
    from sgb200 import _lib
    n, ci, co, h, k, pad, nchw = case
    torch.manual_seed(21)
    torch.backends.cudnn.allow_tf32 = True
    x = torch.randn(n, ci, h, h).to(DEV, dtype)
    if not nchw:
        x = _cl(x)
    x.requires_grad_(True)
    w = (torch.randn(co, ci, k, k) / math.sqrt(ci * k * k)).to(DEV, dtype).requires_grad_(True)
    _lib.profile_start()
    y = cg.conv2d(x, w, padding=pad)
    dy = _cl(torch.randn(y.shape).to(DEV, dtype))        # gradients arrive channels_last in the networks
    dx, dw = torch.autograd.grad(y, [x, w], dy)
    torch.cuda.synchronize()
    kinds = set(_lib.profile_stop().summary())
    assert 'conv_fwd_simt' not in kinds and 'conv_wgrad_simt' not in kinds, kinds      # all three on tensor cores
    xo = x.detach().cpu().float().requires_grad_(True)
    wo = w.detach().cpu().float().requires_grad_(True)
    yo = torch.nn.functional.conv2d(xo, wo, padding=pad)
    dxo, dwo = torch.autograd.grad(yo, [xo, wo], dy.cpu().float())
    assert_close(y, yo, TOL, f'{case} {dtype} y')
    assert_close(dx, dxo, TOL, f'{case} {dtype} dx')
    assert_close(dw, dwo, TOL, f'{case} {dtype} dw')


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, W, pad, stride, tensor_cores)
    (2, 128, 64, 16, 16, 0, 2, True),        # G up path: conv_transpose2d stride 2
    (3, 72, 40, 7, 11, 0, 2, True),          # odd sizes, channel tails
    (2, 64, 64, 12, 12, 1, 1, True),         # stride-1 transposed (dgrad form)
    (2, 10, 6, 9, 9, 0, 2, True),            # channel counts that need zero padding: the scale is applied explicitly
    (2, 64, 32, 8, 8, 0, 2, False),          # SIMT route
])
def test_transposed_conv_in_scale_gradients(dtype, case):
    """conv_transpose2d(x * s[n,c]) with the style scale fused into the op: forward, gradients wrt x, w and s against the
    CPU oracle; the weight gradient takes the scale on its dy operand (descriptor out_scale) on the halo-tile route."""
    from sgb200.ops import conv2d_gradfix as cg
    from sgb200 import _lib
    n, ci, co, h, wd, pad, stride, tc = case
    torch.manual_seed(31)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(n, ci, h, wd).to(DEV, dtype)).requires_grad_(True)
    w = (torch.randn(ci, co, 3, 3) / math.sqrt(ci * 9)).to(DEV, dtype).requires_grad_(True)
    s = (torch.randn(n, ci) + 1).to(DEV, dtype).requires_grad_(True)
    cg.use_tensor_cores = tc
    try:
        _lib.profile_start()
        y = cg.conv_transpose2d(x, w, stride=stride, padding=pad, in_scale=s)
        dy = _cl(torch.randn(y.shape).to(DEV, dtype))
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy)
        torch.cuda.synchronize()
        kinds = set(_lib.profile_stop().summary())
    finally:
        cg.use_tensor_cores = True
    if tc:
        assert 'conv_wgrad_tc' in kinds and 'conv_fwd_simt' not in kinds, kinds
        if ci % 8 == 0 and co % 8 == 0:
            assert 'scale_nc' not in kinds, kinds            # the scale rode inside the kernels
    xo = x.detach().cpu().float().requires_grad_(True)
    wo = w.detach().cpu().float().requires_grad_(True)
    so = s.detach().cpu().float().requires_grad_(True)
    yo = torch.nn.functional.conv_transpose2d(xo * so[:, :, None, None], wo, stride=stride, padding=pad)
    dxo, dwo, dso = torch.autograd.grad(yo, [xo, wo, so], dy.cpu().float())
    tol = TOL if tc else 2e-3
    assert_close(y, yo.detach(), tol, f'{case} {dtype} y')
    assert_close(dx, dxo, tol, f'{case} {dtype} dx')
    assert_close(dw, dwo, tol, f'{case} {dtype} dw')
    assert_close(ds, dso, tol, f'{case} {dtype} ds')


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, W, transposed)
    (3, 64, 3, 32, 32, False),          # toRGB
    (2, 128, 3, 17, 23, False),         # odd sizes: partial last block
    (2, 64, 3, 16, 16, True),           # data gradient of fromRGB
    (2, 32, 8, 9, 9, False),            # the 8-accumulator instantiation
    (1, 512, 1, 4, 4, False),
])
def test_small_co_1x1_kernel(dtype, case):
    """1x1 convolutions with <= 8 output channels run the bandwidth kernel (conv_small.cu), style scale included;
    forward and all three gradients against the CPU oracle."""
    from sgb200.ops import conv2d_gradfix as cg
    from sgb200 import _lib
    n, ci, co, h, wd, tr = case
    torch.manual_seed(41)
    torch.backends.cudnn.allow_tf32 = True
    x = _cl(torch.randn(n, ci, h, wd).to(DEV, dtype)).requires_grad_(True)
    w = (torch.randn((ci, co, 1, 1) if tr else (co, ci, 1, 1)) / math.sqrt(ci)).to(DEV, dtype).requires_grad_(True)
    s = (torch.randn(n, ci) + 1).to(DEV, dtype).requires_grad_(True)
    op = cg.conv_transpose2d if tr else cg.conv2d
    fo = torch.nn.functional.conv_transpose2d if tr else torch.nn.functional.conv2d
    for scale in (None, s):
        _lib.profile_start()
        y = op(x, w, in_scale=scale)
        torch.cuda.synchronize()
        assert set(_lib.profile_stop().summary()) == {'conv_fwd_small'}
        dy = _cl(torch.randn(y.shape).to(DEV, dtype))
        ins = [x, w] + ([scale] if scale is not None else [])
        grads = torch.autograd.grad(y, ins, dy)
        xo = x.detach().cpu().float().requires_grad_(True)
        wo = w.detach().cpu().float().requires_grad_(True)
        so = s.detach().cpu().float().requires_grad_(True)
        yo = fo(xo * so[:, :, None, None] if scale is not None else xo, wo)
        gos = torch.autograd.grad(yo, [xo, wo] + ([so] if scale is not None else []), dy.cpu().float())
        assert_close(y, yo.detach(), TOL, f'{case} {dtype} y scale={scale is not None}')
        for g, go, name in zip(grads, gos, ('dx', 'dw', 'ds')):
            assert_close(g, go, TOL, f'{case} {dtype} {name} scale={scale is not None}')


# ---- second-order gradients THROUGH the tensor-core kernels --------------------------------------------------------
# conv2d_gradfix.py:119-165 of the reference: the backward of a convolution is the other convolution + Conv2dGradWeight, whose own
# backward is two more convolutions.  R1 differentiates |d logits / d image|^2 wrt the weights (dgrad -> its weight gradient);
# a gradient penalty without no_weight_gradients() differentiates dw itself (Conv2dGradWeight.backward).  Oracle: the same
# expressions with torch.nn.functional.conv2d in fp64 on the CPU (what the reference's ops resolve to).
def _second_order_terms(conv, x, w, dy, v_x, v_w):
    y = conv(x, w)
    dx, dw = torch.autograd.grad(y, [x, w], dy, create_graph=True)
    # (i) R1 style: L1 = <dx, dx> ; (ii) full: L2 = <dx, v_x> + <dw, v_w>, differentiated wrt x, w and dy
    L1 = dx.square().sum()
    g1_w, g1_dy = torch.autograd.grad(L1, [w, dy], retain_graph=True)
    L2 = (dx * v_x).sum() + (dw * v_w).sum()
    g2_x, g2_w, g2_dy = torch.autograd.grad(L2, [x, w, dy])
    return dict(y=y, dx=dx, dw=dw, g1_w=g1_w, g1_dy=g1_dy, g2_x=g2_x, g2_w=g2_w, g2_dy=g2_dy)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('case', [
    # (N, Ci, Co, H, k, stride, pad, transposed)
    (4, 64, 64, 32, 3, 1, 1, False),       # stride-1 3x3 (conv1 of every block)
    (4, 128, 64, 16, 3, 2, 0, True),       # transposed stride 2 (G up path)
    (4, 64, 128, 33, 3, 2, 0, False),      # stride 2 after the FIR (D down path)
    (2, 256, 256, 8, 3, 1, 1, False),      # two K blocks of 128, BN = 256
    (4, 64, 64, 16, 1, 1, 0, False),       # 1x1 (skip / fromRGB geometry)
])
def test_tensor_core_double_backward(dtype, case):
    from sgb200.ops import conv2d_gradfix as cg
    from sgb200 import _lib
    n, ci, co, h, k, stride, pad, tr = case
    torch.manual_seed(11)
    torch.backends.cudnn.allow_tf32 = True
    wshape = (ci, co, k, k) if tr else (co, ci, k, k)
    x0 = torch.randn(n, ci, h, h)
    w0 = torch.randn(wshape) / math.sqrt(ci * k * k)
    fo = torch.nn.functional.conv_transpose2d if tr else torch.nn.functional.conv2d
    yo_shape = fo(x0, w0, stride=stride, padding=pad).shape
    dy0, vx0, vw0 = torch.randn(yo_shape), torch.randn(n, ci, h, h), torch.randn(wshape)

    def leaf(t, dev, dt, cl=False):
        t = t.to(dev, dt)
        if cl and t.ndim == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        return t.requires_grad_(True)

    ref = _second_order_terms(lambda a, b: fo(a, b, stride=stride, padding=pad),
                              leaf(x0, 'cpu', torch.float64), leaf(w0, 'cpu', torch.float64), leaf(dy0, 'cpu', torch.float64),
                              vx0.double(), vw0.double())
    op = cg.conv_transpose2d if tr else cg.conv2d
    _lib.profile_start()
    got = _second_order_terms(lambda a, b: op(a, b, stride=stride, padding=pad),
                              leaf(x0, DEV, dtype, True), leaf(w0, DEV, dtype), leaf(dy0, DEV, dtype, True),
                              vx0.to(DEV, dtype).contiguous(memory_format=torch.channels_last), vw0.to(DEV, dtype))
    torch.cuda.synchronize()
    summ = _lib.profile_stop().summary()
    assert 'conv_fwd_simt' not in summ and 'conv_wgrad_simt' not in summ, f'SIMT fallback in the tensor-core test: {sorted(summ)}'
    assert summ['conv_fwd_tc']['launches'] >= 6 and summ['conv_wgrad_tc']['launches'] >= 3, summ
    tol = 1e-2 if dtype == torch.float32 else 2e-2       # fp16 second-order terms: inputs AND first-order results are rounded to fp16
    for key in ref:
        assert_close(got[key].float(), ref[key].float(), tol, f'{case} {dtype} {key}')
