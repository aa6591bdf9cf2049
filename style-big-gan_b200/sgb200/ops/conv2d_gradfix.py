"""conv2d / conv_transpose2d with gradients of arbitrary order, on the libsgb200 convolution kernels.

Public interface of the reference's `stylegan2ada/torch_utils/ops/conv2d_gradfix.py`: `conv2d` (:33),
`conv_transpose2d` (:38), the `no_weight_gradients()` context manager (:25-31) and the module attributes
`enabled` / `weight_gradients_disabled` (:22-23) that `train_parts/regularizations.py:27,48` and
`trainers.py:512` touch.  As in the reference, the backward of a convolution is the *other* convolution
(data gradient) plus a weight-gradient op whose own backward is two more convolutions (:119-165), which is
what makes R1 / path-length double backward work.

Extensions used by our own conv2d_resample / modulated_conv2d (keyword-only, default off):
  flip_weight=False   use the spatially flipped kernel without materialising w.flip([2, 3])
  in_scale=[N,Ci]     multiply x by a per-sample per-channel scale while it is loaded (style modulation)
"""
import contextlib

import torch

from .. import _lib

enabled = True                      # kept for API compatibility; the custom op is the only implementation
use_tensor_cores = True             # False routes every convolution to the SIMT kernels (debugging / A-B tests)
use_halo_kernel = True              # False keeps tensor cores but skips the halo-tile kernels (A-B tests)
halo_gt = 0                         # tiles per super-tile of the halo-tile kernel: 0 = automatic, 1 / 2 / 4 forced (tests, tuning)
channels_last_for_tensor_cores = True   # NCHW operands of a convolution that is otherwise eligible for the tensor-core kernels are
                                    # re-laid out to channels_last (one pass) and the result is returned channels_last -- same
                                    # logical shape, different strides; what cuDNN does internally for its NHWC tensor-core
                                    # kernels.  This is how the reference's unchanged callers (fp32 blocks are NCHW,
                                    # generators.py:392, discriminators.py:247) reach the tcgen05 kernels.  False = such
                                    # operands run the SIMT kernels and keep their layout.
weight_gradients_disabled = False   # forcefully disable computation of gradients with respect to the weights


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, *, flip_weight=False, in_scale=None):
    _lib.require_cuda(input, 'input')
    op = _conv2d_op(transpose=False, weight_shape=weight.shape, stride=stride, padding=padding, output_padding=0,
                    dilation=dilation, groups=groups, flip=bool(flip_weight))
    return op.apply(input, weight, bias, in_scale)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1, *,
                     flip_weight=False, in_scale=None):
    _lib.require_cuda(input, 'input')
    op = _conv2d_op(transpose=True, weight_shape=weight.shape, stride=stride, padding=padding,
                    output_padding=output_padding, dilation=dilation, groups=groups, flip=bool(flip_weight))
    return op.apply(input, weight, bias, in_scale)


def _pair(xs):
    xs = tuple(xs) if isinstance(xs, (tuple, list)) else (xs, xs)
    assert len(xs) == 2 and all(isinstance(x, int) for x in xs)
    return xs


def _make_desc(x, y, transposed, ci, co, kh, kw, stride, pad, groups, flip, in_scale=None, bias=None):
    d = _lib.ConvDesc()
    d.dtype = _lib.dtype_code(x)
    d.transposed = int(transposed)
    d.n, d.ci, d.co = x.shape[0], ci, co
    d.in_h, d.in_w, d.out_h, d.out_w = x.shape[2], x.shape[3], y.shape[2], y.shape[3]
    d.kh, d.kw, d.stride, d.pad_y, d.pad_x, d.groups, d.flip = kh, kw, stride, pad[0], pad[1], groups, int(flip)
    d.x_strides = _lib.strides4(x)
    d.y_strides = _lib.strides4(y)
    d.in_scale = _lib.ptr(in_scale)
    d.out_scale = None
    d.noise = None
    d.bias = _lib.ptr(bias)
    d.act = 1 if bias is not None else 0       # linear, gain 1, no clamp == plain bias add
    d.alpha, d.gain, d.clamp = 0.0, 1.0, -1.0
    # fp32 tensors: TF32 tensor cores only if the caller allows it the way it would for cuDNN
    # (reference trainers.py:511 sets torch.backends.cudnn.allow_tf32 from perf.allow_tf32)
    d.strict_fp32 = 0 if torch.backends.cudnn.allow_tf32 else 1
    d.force_simt = (0 if use_halo_kernel else 2) if use_tensor_cores else 1
    d.halo_gt = int(halo_gt)
    d.workspace, d.workspace_bytes = None, 0
    return d


def _attach_workspace(d, device):
    """Scratch for the tensor-core path (re-packed weights); returns the tensor so it outlives the launch."""
    if not use_tensor_cores:
        return None
    nbytes = int(_lib.lib().sgb_conv2d_workspace_bytes(d))
    if nbytes <= 0:
        return None
    ws = torch.empty([nbytes], dtype=torch.uint8, device=device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    return ws


def _tc_multiple(x, groups):
    """Channel multiple (one 16-byte vector) the tensor-core kernels need, or 0 when they will not be used."""
    if not use_tensor_cores or groups != 1:
        return 0
    if x.dtype in (torch.float16, torch.bfloat16):
        return 8
    if x.dtype == torch.float32 and torch.backends.cudnn.allow_tf32:
        return 4
    return 0


def _pad_channels(t, mult):
    """[N,C,H,W] -> channels_last copy with C padded with zeros to a multiple of `mult` (RGB tensors, the
    minibatch-stddev channel): makes the convolution eligible for the tensor-core kernels."""
    n, c, h, w = t.shape
    cp = (-c) % mult
    out = torch.empty([n, c + cp, h, w], dtype=t.dtype, device=t.device, memory_format=torch.channels_last)
    out[:, :c].copy_(t)
    if cp:
        out[:, c:].zero_()
    return out


def _pad_dim(t, dim, mult):
    cp = (-t.shape[dim]) % mult
    if cp == 0:
        return t
    shape = list(t.shape)
    shape[dim] = cp
    return torch.cat([t, t.new_zeros(shape)], dim=dim)


def _scale_arg(in_scale, x):
    if in_scale is None:
        return None
    assert in_scale.shape == (x.shape[0], x.shape[1])
    return in_scale.detach().to(_lib.acc_dtype(x.dtype)).contiguous()


def _modulation_backward(g, x, in_scale, need_x, need_s):
    """gradients of y = conv(x * s) wrt x and s from g = d(x * s): one pass (sgb_mod_bwd) when no graph is being recorded,
    else the two differentiable ops of fma.py"""
    from . import fma as _fma
    vec = 16 // g.element_size() if g.dtype in (torch.float32, torch.float16, torch.bfloat16) else 0
    c = g.shape[1]
    if (not torch.is_grad_enabled() and vec and c % vec == 0 and c // vec <= 256 and g.numel() > 0
            and _lib.is_channels_last(g) and _lib.is_channels_last(x)):
        n, _, h, w = g.shape
        gx = torch.empty_like(g, memory_format=torch.channels_last) if need_x else None
        gs = torch.empty([n, c], dtype=torch.float32, device=g.device) if need_s else None
        s32 = in_scale.detach().to(torch.float32).contiguous()
        with torch.cuda.device(g.device), _lib.prof('mod_bwd', 0.0, 3 * g.numel() * g.element_size()):
            rc = _lib.lib().sgb_mod_bwd(_lib.ptr(g), _lib.ptr(x), _lib.ptr(s32), _lib.ptr(gx), _lib.ptr(gs), _lib.dtype_code(g),
                                        n, c, h * w, _lib.stream_ptr(g.device))
        _lib.check(rc, 'mod_bwd')
        return gx, (gs.to(in_scale.dtype) if gs is not None else None)
    gx = _fma.scale_nc(g, in_scale) if need_x else None
    gs = _fma.mul_sum_hw(g, x).to(in_scale.dtype) if need_s else None
    return gx, gs


_cache = dict()


def _conv2d_op(transpose, weight_shape, stride, padding, output_padding, dilation, groups, flip):
    weight_shape = tuple(int(s) for s in weight_shape)
    stride = _pair(stride)
    padding = _pair(padding)
    output_padding = _pair(output_padding)
    dilation = _pair(dilation)
    key = (transpose, weight_shape, stride, padding, output_padding, dilation, groups, flip)
    if key in _cache:
        return _cache[key]

    assert groups >= 1 and len(weight_shape) == 4
    if stride[0] != stride[1]:
        raise NotImplementedError('sgb200 conv: stride must be the same in both directions')
    if dilation != (1, 1):
        raise NotImplementedError('sgb200 conv: dilation is not supported')
    assert all(p >= 0 for p in padding)
    if not transpose:
        assert output_padding == (0, 0)
    else:
        assert all(0 <= output_padding[i] < max(stride[i], dilation[i]) for i in range(2))
    s = stride[0]
    kh, kw = weight_shape[2], weight_shape[3]
    # channel counts of the op's input / output
    if not transpose:
        co, ci = weight_shape[0], weight_shape[1] * groups
    else:
        ci, co = weight_shape[0], weight_shape[1] * groups

    def out_hw(ih, iw):
        if not transpose:
            return (ih + 2 * padding[0] - kh) // s + 1, (iw + 2 * padding[1] - kw) // s + 1
        return (ih - 1) * s - 2 * padding[0] + kh + output_padding[0], (iw - 1) * s - 2 * padding[1] + kw + output_padding[1]

    def calc_output_padding(input_shape, output_shape):
        if transpose:
            return [0, 0]
        return [input_shape[i + 2] - (output_shape[i + 2] - 1) * s - (1 - 2 * padding[i]) - (weight_shape[i + 2] - 1)
                for i in range(2)]

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias, in_scale):
            assert tuple(weight.shape) == weight_shape
            if input.ndim != 4 or input.shape[1] != ci:
                raise RuntimeError(f'conv: expected input [N, {ci}, H, W], got {tuple(input.shape)}')
            if weight.dtype != input.dtype:
                raise RuntimeError('conv: weight and input must have the same dtype')
            w = weight.contiguous()
            oh, ow = out_hw(input.shape[2], input.shape[3])
            if oh < 1 or ow < 1:
                raise RuntimeError('conv: output would be empty')
            sc = _scale_arg(in_scale, input)
            x_, ci_, fmt = input, ci, _lib.out_format(input)
            mult = _tc_multiple(input, groups)
            if mult and ci % mult != 0 and input.numel() > 0:
                # few / odd input channels (RGB, minibatch-stddev): zero-pad them to one 16-byte vector
                x_ = _pad_channels(input, mult)
                w = _pad_dim(w, 0 if transpose else 1, mult)
                sc = _pad_dim(sc, 1, mult) if sc is not None else None
                ci_, fmt = x_.shape[1], torch.channels_last
            elif mult and channels_last_for_tensor_cores and input.numel() > 0 and not _lib.is_channels_last(input):
                x_, fmt = input.contiguous(memory_format=torch.channels_last), torch.channels_last
            y = torch.empty([input.shape[0], co, oh, ow], dtype=input.dtype, device=input.device, memory_format=fmt)
            b = bias.contiguous() if bias is not None else None
            if y.numel() > 0:
                d = _make_desc(x_, y, transpose, ci_, co, kh, kw, s, padding, groups, flip, sc, b)
                # algorithmic work: non-zero MACs only for the transposed form
                px = (input.shape[2] * input.shape[3]) if transpose else (oh * ow)
                flops = 2.0 * input.shape[0] * px * (co // groups) * ci * kh * kw
                nbytes = (input.numel() + y.numel() + w.numel()) * input.element_size()
                ws = _attach_workspace(d, input.device)
                tc = _lib.lib().sgb_conv2d_uses_tensor_cores(d)
                tag = (f"{str(input.dtype)[6:]} x[{input.shape[0]},{ci},{input.shape[2]},{input.shape[3]}] co{co} k{kh} s{s}"
                       f"{' T' if transpose else ''}{' mod' if sc is not None else ''}{' tma' if tc == 3 else ''}") if _lib.PROFILE is not None else None
                with torch.cuda.device(input.device), _lib.prof(('conv_fwd_simt', 'conv_fwd_tc', 'conv_fwd_small', 'conv_fwd_tc')[tc], flops, nbytes, tag):
                    rc = _lib.lib().sgb_conv2d_forward(d, _lib.ptr(x_), _lib.ptr(w), _lib.ptr(y), _lib.stream_ptr(input.device))
                _lib.check(rc, 'conv2d_forward')
            ctx.save_for_backward(input, weight, in_scale)
            ctx.has_bias = bias is not None
            return y

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, in_scale = ctx.saved_tensors
            grad_input = grad_weight = grad_bias = grad_scale = None
            need_x = ctx.needs_input_grad[0]
            need_s = in_scale is not None and ctx.needs_input_grad[3]
            if need_x or need_s:
                p = calc_output_padding(input.shape, grad_output.shape)
                dgrad = _conv2d_op(transpose=(not transpose), weight_shape=weight_shape, stride=stride, padding=padding,
                                   output_padding=p, dilation=dilation, groups=groups, flip=flip)
                g = dgrad.apply(grad_output, weight, None, None)        # gradient wrt (input * in_scale)
                assert g.shape == input.shape
                if in_scale is None:
                    grad_input = g
                else:
                    grad_input, grad_scale = _modulation_backward(g, input, in_scale, need_x, need_s)
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                grad_weight = Conv2dGradWeight.apply(grad_output, input, in_scale)
                assert tuple(grad_weight.shape) == weight_shape
            if ctx.has_bias and ctx.needs_input_grad[2]:
                grad_bias = grad_output.sum([0, 2, 3])
            return grad_input, grad_weight, grad_bias, grad_scale

    class Conv2dGradWeight(torch.autograd.Function):
        """dw of the op above.  For a transposed op the roles of (input, grad_output) swap: conv_transpose2d
        is the adjoint of the conv2d that maps the op's output space to its input space."""
        @staticmethod
        def forward(ctx, grad_output, input, in_scale):
            go_, in_ = grad_output, input
            if (_tc_multiple(input, groups) and channels_last_for_tensor_cores and input.numel() > 0 and grad_output.numel() > 0
                    and ci % _tc_multiple(input, groups) == 0 and co % _tc_multiple(input, groups) == 0):
                grad_output = grad_output.contiguous(memory_format=torch.channels_last)
                input = input.contiguous(memory_format=torch.channels_last)
            if not transpose:
                x_, dy_, sc = input, grad_output, _scale_arg(in_scale, input)
                d = _make_desc(x_, dy_, False, ci, co, kh, kw, s, padding, groups, flip, sc)
            else:
                # the op's input plays the role of dy: its style scale goes into the descriptor's out_scale (taken by
                # the halo-tile kernels only; otherwise the input is scaled explicitly first)
                x_, dy_, sc = grad_output, input, None
                d = _make_desc(x_, dy_, False, co, ci, kh, kw, s, padding, groups, flip)
                if in_scale is not None:
                    mult_ = _tc_multiple(input, groups)
                    if _lib.lib().sgb_conv2d_wgrad_uses_tensor_cores(d) == 2 and not (mult_ and (ci % mult_ or co % mult_)):
                        sc_dy = _scale_arg(in_scale, input)
                        d.out_scale = _lib.ptr(sc_dy)
                    else:
                        from . import fma as _fma
                        dy_ = _fma.scale_nc(input.detach(), in_scale.detach())
                        d = _make_desc(x_, dy_, False, co, ci, kh, kw, s, padding, groups, flip)
            flops = 2.0 * dy_.shape[0] * dy_.shape[2] * dy_.shape[3] * dy_.shape[1] * (x_.shape[1] // groups) * kh * kw
            nbytes = (x_.numel() + dy_.numel()) * x_.element_size()
            # dw has the layout of the weight of the NON-transposed conv x_ -> dy_: [C(dy_), C(x_)/groups, kh, kw]
            mult = _tc_multiple(input, groups)
            cx, cy = x_.shape[1], dy_.shape[1]
            padded = bool(mult) and (cx % mult != 0 or cy % mult != 0) and x_.numel() > 0 and dy_.numel() > 0
            if padded:      # few / odd channels: zero-pad to one 16-byte vector so the tensor-core kernel applies
                if cx % mult != 0:
                    x_ = _pad_channels(x_, mult)
                    if d.in_scale:
                        sc = _pad_dim(sc, 1, mult)
                if cy % mult != 0:
                    dy_ = _pad_channels(dy_, mult)
                x_ = x_.contiguous(memory_format=torch.channels_last)
                dy_ = dy_.contiguous(memory_format=torch.channels_last)
                d = _make_desc(x_, dy_, False, x_.shape[1], dy_.shape[1], kh, kw, s, padding, groups, flip,
                               sc if d.in_scale else None)
            dwf = torch.empty([dy_.shape[1], x_.shape[1] // groups, kh, kw], dtype=_lib.acc_dtype(input.dtype), device=input.device)
            nbytes += dwf.numel() * dwf.element_size()
            tc = _lib.lib().sgb_conv2d_wgrad_uses_tensor_cores(d)
            tag = (f"{str(input.dtype)[6:]} x[{x_.shape[0]},{x_.shape[1]},{x_.shape[2]},{x_.shape[3]}] dy[{dy_.shape[1]},{dy_.shape[2]},"
                   f"{dy_.shape[3]}] k{kh} s{s}") if _lib.PROFILE is not None else None
            with torch.cuda.device(input.device), _lib.prof('conv_wgrad_tc' if tc else 'conv_wgrad_simt', flops, nbytes, tag):
                rc = _lib.lib().sgb_conv2d_wgrad(d, _lib.ptr(x_), _lib.ptr(dy_), _lib.ptr(dwf), _lib.stream_ptr(input.device))
            _lib.check(rc, 'conv2d_wgrad')
            if padded:
                dwf = dwf[:cy, :cx // groups]
            # the non-transposed layout [C(dy_), C(x_), kh, kw] is the op's weight layout in both cases:
            # conv2d: [co, ci]; conv_transpose2d (x_ = grad_output, dy_ = input): [ci, co]
            dw = dwf.reshape(weight_shape) if not padded else dwf.contiguous().reshape(weight_shape)
            ctx.save_for_backward(go_, in_, in_scale)
            return dw.to(input.dtype)

        @staticmethod
        def backward(ctx, grad2_grad_weight):
            grad_output, input, in_scale = ctx.saved_tensors
            grad2_grad_output = grad2_input = grad2_scale = None
            ddw = grad2_grad_weight.to(input.dtype)
            if ctx.needs_input_grad[0]:
                grad2_grad_output = Conv2d.apply(input, ddw, None, in_scale)
                assert grad2_grad_output.shape == grad_output.shape
            need_x = ctx.needs_input_grad[1]
            need_s = in_scale is not None and ctx.needs_input_grad[2]
            if need_x or need_s:
                p = calc_output_padding(input.shape, grad_output.shape)
                dgrad = _conv2d_op(transpose=(not transpose), weight_shape=weight_shape, stride=stride, padding=padding,
                                   output_padding=p, dilation=dilation, groups=groups, flip=flip)
                g = dgrad.apply(grad_output, ddw, None, None)
                assert g.shape == input.shape
                if in_scale is None:
                    grad2_input = g
                else:
                    from . import fma as _fma
                    if need_x:
                        grad2_input = _fma.scale_nc(g, in_scale)
                    if need_s:
                        grad2_scale = _fma.mul_sum_hw(g, input).to(in_scale.dtype)
            return grad2_grad_output, grad2_input, grad2_scale

    Conv2d.grad_weight_op = Conv2dGradWeight
    Conv2d.calc_output_padding = staticmethod(calc_output_padding)
    _cache[key] = Conv2d
    return Conv2d
