"""conv2d + the epilogue every StyleGAN2 layer runs after it, as ONE kernel launch forward and (first order) ONE
activation-sized pass backward:

    y = bias_act( conv2d(x * styles[n,ci], w, stride, padding) * dcoefs[n,co] + noise[n,1,h,w],  b, act, gain, clamp )

This is what `SynthesisLayer.forward` (generators.py:310-329: modulated_conv2d -> fma -> bias_act) and
`Conv2dLayer.forward` (discriminators.py:115-124: conv2d_resample -> bias_act) compute; the reference spends one
activation-sized HBM round trip per arrow.  Here the demodulation scale, the noise, the bias, lrelu|linear, the gain
and the clamp run in the tcgen05 convolution's epilogue while the accumulator leaves TMEM (`sgb_conv_desc.out_scale /
noise / bias / act`), and the backward of that epilogue is a single pass (`sgb_fused_epilogue_bwd`) that re-derives
the pre-activation from y.

Gradients of higher order: when the backward itself is being recorded (`create_graph=True`: R1 / path-length
regularisation) it is evaluated through the un-fused differentiable ops (conv2d_gradfix, fma, bias_act), which is
exactly the reference's composition.  `conv2d_gradfix.no_weight_gradients()` is honoured on both routes.
"""
import os

import torch

from .. import _lib
from . import bias_act as _ba
from . import conv2d_gradfix as _cg
from . import fma as _fma

_ACT_ID = {'linear': 1, 'lrelu': 3}
# Which convolutions run the tail in their epilogue (SGB_FUSED_CONV):
#   '0' (default)  none: every convolution is followed by the ONE-pass tail kernel (sgb_scale_bias_act / bias_act)
#   'tma'          those that take the TMA-staged kernel (csrc/conv_tma.cuh), whose epilogue stages the per-channel vectors of a
#                  tile in shared memory once
#   '1'            every tensor-core convolution (tests, A/B)
# Measured on a B200 (profiles/README.md): the fused forward is SLOWER than the pass it removes in both kernels -- round 1,
# conv_halo_kernel (one __ldg per element in the epilogue): 93.8 vs 78.7 ms/step; round 2, conv_tma_kernel (vectors staged in
# shared memory): ffhq256 76.6 vs 73.0 ms/step, config-f 1024 64.0 vs 59.3.  With the MMA issue cost cut 3x (warp-uniform
# issue) the four epilogue warps are the critical path of the few-channel layers, and every instruction added to them shows.
enabled = os.environ.get('SGB_FUSED_CONV', '0')
if enabled in ('0', '1'):
    enabled = enabled == '1'
# SGB_FUSED_CONV_MAIN ('0' | 'tma' | '1'): the same switch, but only for phases whose backward is first order (Gmain / Dmain);
# the training harness sets `enabled` from it per phase.  Under create_graph=True (Greg / Dreg) the fused forward is a loss
# whatever the kernel costs: its differentiable backward re-evaluates the convolution (the un-fused tail re-evaluates only
# itself, from the saved convolution output).
main_phase_mode = os.environ.get('SGB_FUSED_CONV_MAIN', '0')
if main_phase_mode in ('0', '1'):
    main_phase_mode = main_phase_mode == '1'


class phase_mode:
    """with fused_conv.phase_mode(first_order=True|False): ... -- applies `main_phase_mode` for the duration of a training phase"""
    def __init__(self, first_order):
        self.first_order = first_order

    def __enter__(self):
        global enabled
        self.old = enabled
        if main_phase_mode:
            enabled = main_phase_mode if self.first_order else False

    def __exit__(self, *a):
        global enabled
        enabled = self.old


def _noise4(noise, n, h, w):
    if noise is None:
        return None
    if noise.ndim < 4 or noise.shape[0] != n:
        noise = noise.expand(n, 1, h, w)
    return noise


def _conv_then_tail(x, w, b, styles, dcoefs, noise, stride, padding, flip_weight, act, alpha, gain, clamp):
    """convolution followed by the one-pass tail (differentiable; the default route of layers that are not fused)"""
    y = _cg.conv2d(x, w, stride=stride, padding=padding, flip_weight=(not flip_weight), in_scale=styles)
    if dcoefs is None and noise is None:
        return _ba.bias_act(y, b, act=act, alpha=alpha, gain=gain, clamp=clamp)
    return scale_bias_act(y, b, dcoefs, noise, act=act, alpha=alpha, gain=gain, clamp=clamp)


def _unfused(x, w, b, styles, dcoefs, noise, stride, padding, flip_weight, act, alpha, gain, clamp):
    y = _cg.conv2d(x, w, stride=stride, padding=padding, flip_weight=(not flip_weight), in_scale=styles)
    if noise is not None:
        noise = _noise4(noise, y.shape[0], y.shape[2], y.shape[3]).to(y.dtype)
    if dcoefs is not None:
        y = _fma.scale_nc(y, dcoefs, noise)
    elif noise is not None:
        y = y + noise
    return _ba.bias_act(y, b, act=act, alpha=alpha, gain=gain, clamp=clamp)


def _fusable(x, w, b, styles, dcoefs, noise, act, gain=1.0):
    if not enabled or gain == 0 or act not in _ACT_ID or not x.is_cuda or x.ndim != 4 or x.numel() == 0:
        return False
    mult = _cg._tc_multiple(x, 1)
    if not mult or x.shape[1] % mult != 0:
        return False
    if not _lib.is_channels_last(x):
        return False
    return True


def conv2d_bias_act(x, w, b=None, *, stride=1, padding=0, flip_weight=True, styles=None, dcoefs=None, noise=None,
                    act='linear', alpha=None, gain=None, clamp=None):
    """See the module docstring.  `flip_weight=True` means correlation (the weight as stored), like conv2d_resample."""
    spec = _ba.activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    padding = tuple(padding) if isinstance(padding, (tuple, list)) else (padding, padding)
    if not _fusable(x, w, b, styles, dcoefs, noise, act, gain):
        return _conv_then_tail(x, w, b, styles, dcoefs, noise, stride, padding, flip_weight, act, alpha, gain, clamp if clamp >= 0 else None)
    op = _fused_op(tuple(int(s) for s in w.shape), int(stride), padding, bool(flip_weight), act, alpha, gain, clamp)
    return op.apply(x, w, b, styles, dcoefs, noise)


_cache = dict()


def _fused_op(weight_shape, stride, padding, flip_weight, act, alpha, gain, clamp):
    key = (weight_shape, stride, padding, flip_weight, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]
    co, ci, kh, kw = weight_shape
    flip = not flip_weight
    conv = _cg._conv2d_op(transpose=False, weight_shape=weight_shape, stride=stride, padding=padding, output_padding=0,
                          dilation=1, groups=1, flip=flip)
    clamp_arg = clamp if clamp >= 0 else None

    class FusedConvBiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, styles, dcoefs, noise):
            n = x.shape[0]
            oh = (x.shape[2] + 2 * padding[0] - kh) // stride + 1
            ow = (x.shape[3] + 2 * padding[1] - kw) // stride + 1
            if oh < 1 or ow < 1:
                raise RuntimeError('conv: output would be empty')
            if w.dtype != x.dtype:
                raise RuntimeError('conv: weight and input must have the same dtype')
            y = torch.empty([n, co, oh, ow], dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
            wc = w.contiguous()
            sc = _cg._scale_arg(styles, x)
            d = _cg._make_desc(x, y, False, ci, co, kh, kw, stride, padding, 1, flip, sc, None)
            acc = _lib.acc_dtype(x.dtype)
            dc = dcoefs.detach().reshape(n, co).to(acc).contiguous() if dcoefs is not None else None
            nz = _noise4(noise, n, oh, ow).detach().to(acc).reshape(n, oh, ow).contiguous() if noise is not None else None
            bb = b.detach().to(x.dtype).contiguous() if b is not None else None
            d.out_scale, d.noise, d.bias = _lib.ptr(dc), _lib.ptr(nz), _lib.ptr(bb)
            d.act, d.alpha, d.gain, d.clamp = _ACT_ID[act], alpha, gain, clamp
            flops = 2.0 * n * oh * ow * co * ci * kh * kw
            nbytes = (x.numel() + y.numel() + wc.numel()) * x.element_size()
            ws = _cg._attach_workspace(d, x.device)
            tc = _lib.lib().sgb_conv2d_uses_tensor_cores(d)
            if enabled == 'tma' and tc != 3:
                # not a TMA-kernel convolution: same values from the convolution + the one-pass tail; the backward below (one
                # pass over dy and y, then the two gradient convolutions) is the same either way
                del ws, y
                with torch.no_grad():
                    y = _conv_then_tail(x, w, b, styles, dcoefs, noise, stride, padding, flip_weight, act, alpha, gain, clamp_arg)
            else:
                tag = (f"{str(x.dtype)[6:]} x[{n},{ci},{x.shape[2]},{x.shape[3]}] co{co} k{kh} s{stride}{' mod' if sc is not None else ''}"
                       f"{' tma' if tc == 3 else ''} +tail") if _lib.PROFILE is not None else None
                with torch.cuda.device(x.device), _lib.prof('conv_fwd_tc' if tc else 'conv_fwd_simt', flops, nbytes, tag):
                    rc = _lib.lib().sgb_conv2d_forward(d, _lib.ptr(x), _lib.ptr(wc), _lib.ptr(y), _lib.stream_ptr(x.device))
                _lib.check(rc, 'conv2d_forward (fused epilogue)')
                del ws
            # y is only kept when the backward depends on it; like the reference (bias_act.py:152-155) a plain linear
            # epilogue does not, so callers may modify its output in place (`y.add_(x)`, discriminators.py:300)
            keep_y = act != 'linear' or clamp >= 0 or dcoefs is not None
            ctx.save_for_backward(x, w, b, styles, dcoefs, noise, y if keep_y else None)
            ctx.out_shape = tuple(y.shape)
            return y

        @staticmethod
        def backward(ctx, dy):
            x, w, b, styles, dcoefs, noise, y = ctx.saved_tensors
            need = ctx.needs_input_grad
            vec = 16 // x.element_size()
            fast = (not torch.is_grad_enabled()) and co % vec == 0 and co // vec <= 256
            if not fast:
                # differentiable route (create_graph=True, or a channel count the one-pass kernel does not take):
                # the reference's own composition, evaluated on the saved inputs
                with torch.enable_grad():
                    y2 = _unfused(x, w, b, styles, dcoefs, noise, stride, padding, flip_weight, act, alpha, gain, clamp_arg)
                    ins = [t for t, nd in zip((x, w, b, styles, dcoefs, noise), need) if nd and t is not None]
                    gs = torch.autograd.grad([y2], ins, [dy], create_graph=torch.is_grad_enabled(), allow_unused=True) if ins else []
                it = iter(gs)
                return tuple(next(it) if (nd and t is not None) else None
                             for t, nd in zip((x, w, b, styles, dcoefs, noise), need))
            n, _, oh, ow = ctx.out_shape
            dy = dy.contiguous(memory_format=torch.channels_last)
            dconv = torch.empty_like(dy, memory_format=torch.channels_last)
            f32 = torch.float32
            dev = dy.device
            db = torch.empty([co], dtype=f32, device=dev) if (b is not None and need[2]) else None
            dsc = torch.empty([n, co], dtype=f32, device=dev) if (dcoefs is not None and need[4]) else None
            dnz = torch.empty([n, oh, ow], dtype=f32, device=dev) if (noise is not None and need[5]) else None
            dc = dcoefs.detach().reshape(n, co).to(f32).contiguous() if dcoefs is not None else None
            nz = _noise4(noise, n, oh, ow).detach().to(f32).reshape(n, oh, ow).contiguous() if noise is not None else None
            bb = b.detach().to(dy.dtype).contiguous() if b is not None else None
            with torch.cuda.device(dev), _lib.prof('fused_epilogue_bwd', 0.0, 3 * dy.numel() * dy.element_size()):
                rc = _lib.lib().sgb_fused_epilogue_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dconv), _lib.ptr(bb), _lib.ptr(dc),
                                                       _lib.ptr(nz), _lib.ptr(db), _lib.ptr(dnz), _lib.ptr(dsc),
                                                       _lib.dtype_code(dy), n, co, oh * ow, _ACT_ID[act], alpha, gain, clamp,
                                                       _lib.stream_ptr(dev))
            _lib.check(rc, 'fused_epilogue_bwd')
            gx = gw = gs_ = None
            need_x, need_s = need[0], styles is not None and need[3]
            if need_x or need_s:
                pad_out = conv.calc_output_padding(x.shape, dconv.shape)
                dgrad = _cg._conv2d_op(transpose=True, weight_shape=weight_shape, stride=stride, padding=padding,
                                       output_padding=pad_out, dilation=1, groups=1, flip=flip)
                g = dgrad.apply(dconv, w, None, None)
                if styles is None:
                    gx = g
                else:
                    if need_x:
                        gx = _fma.scale_nc(g, styles)
                    if need_s:
                        gs_ = _fma.mul_sum_hw(g, x).to(styles.dtype)
            if need[1] and not _cg.weight_gradients_disabled:
                gw = conv.grad_weight_op.apply(dconv, x, styles)
            gb = db.to(b.dtype) if db is not None else None
            gd = dsc.reshape(dcoefs.shape).to(dcoefs.dtype) if dsc is not None else None
            gn = None
            if dnz is not None:
                gn = dnz.reshape(n, 1, oh, ow).to(noise.dtype)
                if tuple(noise.shape) != (n, 1, oh, ow):       # broadcast noise (noise_mode='const'): un-broadcast
                    gn = gn.sum_to_size(noise.shape) if noise.ndim == 4 else gn.sum(dim=0).reshape(noise.shape)
            return gx, gw, gb, gs_, gd, gn

    _cache[key] = FusedConvBiasAct
    return FusedConvBiasAct


# ---------------------------------------------------------------------------------------------------------------------
# The same epilogue as ONE stand-alone pass after an un-fused convolution (or after the FIR pass of an up-sampling layer):
#   y = bias_act(x * dcoefs[n,c] + noise, b, act, gain, clamp)        forward: sgb_scale_bias_act (1 pass instead of 2)
#   backward (first order): sgb_fused_epilogue_bwd (1 pass instead of bias_act-grad, bias sum, dz * dcoefs, sum(dz * x), sum_c(dz))
#   backward under create_graph: the reference's composition (fma -> bias_act) on the saved inputs.
tail_enabled = True
_tail_cache = dict()


def scale_bias_act(x, b, dcoefs, noise, act='linear', alpha=None, gain=None, clamp=None):
    spec = _ba.activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clampf = float(clamp if clamp is not None else -1)
    vec = 16 // x.element_size() if x.dtype in (torch.float32, torch.float16, torch.bfloat16) else 0
    ok = (tail_enabled and act in _ACT_ID and x.is_cuda and x.ndim == 4 and x.numel() > 0 and vec and x.shape[1] % vec == 0
          and x.shape[1] // vec <= 256 and _lib.is_channels_last(x) and gain != 0)
    if not ok:
        return _tail_unfused(x, b, dcoefs, noise, act, alpha, gain, clamp)
    key = (act, alpha, gain, clampf)
    if key not in _tail_cache:
        _tail_cache[key] = _make_tail(act, alpha, gain, clampf)
    return _tail_cache[key].apply(x, b, dcoefs, noise)


def _tail_unfused(x, b, dcoefs, noise, act, alpha, gain, clamp):
    if noise is not None:
        noise = _noise4(noise, x.shape[0], x.shape[2], x.shape[3]).to(x.dtype)
    if dcoefs is not None:
        x = _fma.scale_nc(x, dcoefs, noise)
    elif noise is not None:
        x = x + noise
    return _ba.bias_act(x, b, act=act, alpha=alpha, gain=gain, clamp=clamp)


def _make_tail(act, alpha, gain, clamp):
    clamp_arg = clamp if clamp >= 0 else None

    class ScaleBiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b, dcoefs, noise):
            n, c, h, w = x.shape
            y = torch.empty_like(x, memory_format=torch.channels_last)
            f32 = torch.float32
            dc = dcoefs.detach().reshape(n, c).to(f32).contiguous() if dcoefs is not None else None
            nz = _noise4(noise, n, h, w).detach().to(f32).reshape(n, h, w).contiguous() if noise is not None else None
            bb = b.detach().to(x.dtype).contiguous() if b is not None else None
            with torch.cuda.device(x.device), _lib.prof('scale_bias_act', 0.0, 2 * x.numel() * x.element_size()):
                rc = _lib.lib().sgb_scale_bias_act(_lib.ptr(x), _lib.ptr(bb), _lib.ptr(dc), _lib.ptr(nz), _lib.ptr(y), _lib.dtype_code(x),
                                                   n, c, h * w, _ACT_ID[act], alpha, gain, clamp, _lib.stream_ptr(x.device))
            _lib.check(rc, 'scale_bias_act')
            # x is kept only for the differentiable route (and costs no extra pass); y serves the one-pass backward
            ctx.save_for_backward(x, b, dcoefs, noise, y)
            return y

        @staticmethod
        def backward(ctx, dy):
            x, b, dcoefs, noise, y = ctx.saved_tensors
            need = ctx.needs_input_grad
            if torch.is_grad_enabled():
                with torch.enable_grad():
                    y2 = _tail_unfused(x, b, dcoefs, noise, act, alpha, gain, clamp_arg)
                    ins = [t for t, nd in zip((x, b, dcoefs, noise), need) if nd and t is not None]
                    gs = torch.autograd.grad([y2], ins, [dy], create_graph=True, allow_unused=True) if ins else []
                it = iter(gs)
                return tuple(next(it) if (nd and t is not None) else None for t, nd in zip((x, b, dcoefs, noise), need))
            n, c, h, w = y.shape
            dy = dy.contiguous(memory_format=torch.channels_last)
            dx = torch.empty_like(y, memory_format=torch.channels_last)
            f32, dev = torch.float32, y.device
            db = torch.empty([c], dtype=f32, device=dev) if (b is not None and need[1]) else None
            dsc = torch.empty([n, c], dtype=f32, device=dev) if (dcoefs is not None and need[2]) else None
            dnz = torch.empty([n, h, w], dtype=f32, device=dev) if (noise is not None and need[3]) else None
            dc = dcoefs.detach().reshape(n, c).to(f32).contiguous() if dcoefs is not None else None
            nz = _noise4(noise, n, h, w).detach().to(f32).reshape(n, h, w).contiguous() if noise is not None else None
            bb = b.detach().to(y.dtype).contiguous() if b is not None else None
            with torch.cuda.device(dev), _lib.prof('fused_epilogue_bwd', 0.0, 3 * y.numel() * y.element_size()):
                rc = _lib.lib().sgb_fused_epilogue_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), _lib.ptr(bb), _lib.ptr(dc), _lib.ptr(nz),
                                                       _lib.ptr(db), _lib.ptr(dnz), _lib.ptr(dsc), _lib.dtype_code(y), n, c, h * w,
                                                       _ACT_ID[act], alpha, gain, clamp, _lib.stream_ptr(dev))
            _lib.check(rc, 'fused_epilogue_bwd')
            gb = db.to(b.dtype) if db is not None else None
            gd = dsc.reshape(dcoefs.shape).to(dcoefs.dtype) if dsc is not None else None
            gn = None
            if dnz is not None:
                gn = dnz.reshape(n, 1, h, w).to(noise.dtype)
                if tuple(noise.shape) != (n, 1, h, w):
                    gn = gn.sum_to_size(noise.shape) if noise.ndim == 4 else gn.sum(dim=0).reshape(noise.shape)
            return (dx if need[0] else None), gb, gd, gn

    return ScaleBiasAct
