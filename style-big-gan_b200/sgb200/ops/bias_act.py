"""bias_act: fused bias + activation + gain + clamp.

Same public interface as the reference's `stylegan2ada/torch_utils/ops/bias_act.py`
(`bias_act(x, b, dim, act, alpha, gain, clamp, impl)` at :55, table `activation_funcs` at :23-33 whose
`def_gain` the layers read).  Forward, first and second order gradients run on the sm_100a kernels behind
`sgb_bias_act` / `sgb_sum_to_channel`; there is no PyTorch or CPU path.
"""
import math
from types import SimpleNamespace

import torch

from .. import _lib


class _Spec(SimpleNamespace):
    def __getitem__(self, k):          # EasyDict-style access used by some callers
        return getattr(self, k)


# name -> defaults, kernel id (= cuda_idx), which tensors the derivative needs (bias_act.py:23-33)
activation_funcs = {
    'linear':   _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     _Spec(def_alpha=0.0, def_gain=math.sqrt(2.0), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    _Spec(def_alpha=0.2, def_gain=math.sqrt(2.0), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': _Spec(def_alpha=0.0, def_gain=1.0,            cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    _Spec(def_alpha=0.0, def_gain=math.sqrt(2.0), cuda_idx=9, ref='x', has_2nd_grad=True),
}


def _dense_like(x):
    """x in a dense layout the kernel can index flat (NCHW-contiguous or channels_last)."""
    if x.dim() == 4 and _lib.is_channels_last(x):
        return x, torch.channels_last
    return x.contiguous(), torch.contiguous_format


def _launch(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    """One sgb_bias_act call; every tensor shares x's dense layout."""
    y = torch.empty_like(x)     # preserves x's (dense) strides
    if x.numel() == 0:
        return y
    step_b = x.stride(dim) if b is not None else 1
    size_b = b.numel() if b is not None else 1
    nb = x.numel() * x.element_size() * (2 + (xref is not None) + (yref is not None) + (dy is not None))
    with torch.cuda.device(x.device), _lib.prof('bias_act_g%d' % grad, 0.0, nb):
        rc = _lib.lib().sgb_bias_act(_lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), _lib.ptr(y),
                                     _lib.dtype_code(x), grad, spec.cuda_idx, alpha, gain, clamp,
                                     x.numel(), size_b, step_b, _lib.stream_ptr(x.device))
    _lib.check(rc, 'bias_act')
    return y


def _sum_to_bias(dx, dim):
    """db = dx summed over every dimension but `dim` (sgb_sum_to_channel)."""
    c = dx.shape[dim]
    out = torch.empty([c], dtype=_lib.acc_dtype(dx.dtype), device=dx.device)
    if c == 0:
        return out.to(dx.dtype)
    inner = dx.stride(dim)
    outer = dx.numel() // (c * inner) if dx.numel() else 0
    with torch.cuda.device(dx.device), _lib.prof('sum_to_channel', 0.0, dx.numel() * dx.element_size()):
        rc = _lib.lib().sgb_sum_to_channel(_lib.ptr(dx), _lib.ptr(out), _lib.dtype_code(dx), outer, c, inner,
                                           _lib.stream_ptr(dx.device))
    _lib.check(rc, 'sum_to_channel')
    return out.to(dx.dtype)


_cache = dict()
fused_backward = True      # first-order backward: dx and db in one pass (False = the two reference-shaped kernels)


def _bias_act_cuda(dim, act, alpha, gain, clamp):
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]

    # the derivative needs y for the sign (relu family) and for the clamp mask; swish needs x
    need_y = ('y' in spec.ref) or (clamp >= 0 and 'x' not in spec.ref)
    need_x = ('x' in spec.ref) or spec.has_2nd_grad
    trivial = (act == 'linear' and gain == 1 and clamp < 0)

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            x, ctx.fmt = _dense_like(x)
            if b is not None:
                b = b.contiguous()
            if trivial and b is None:
                y = x.clone()          # callers add_ into the result: it must not alias the input
            else:
                y = _launch(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            ctx.has_b = b is not None
            ctx.save_for_backward(x if need_x else None, b if (need_x and b is not None) else None, y if need_y else None)
            return y

        @staticmethod
        def backward(ctx, dy):
            x, b, y = ctx.saved_tensors
            dx = db = None
            # first-order fast path (no graph being recorded): dx and db from ONE pass over (dy, y) instead of the
            # activation-gradient kernel followed by a second read of dx for the bias sum (bias_act.py:172-173)
            if (fused_backward and not trivial and not torch.is_grad_enabled() and ctx.has_b and ctx.needs_input_grad[1]
                    and act in ('linear', 'lrelu') and dim == 1 and dy.dim() == 4 and dy.numel() > 0
                    and dy.dtype in (torch.float32, torch.float16, torch.bfloat16)
                    and (y is None or _lib.is_channels_last(y)) and (y is not None or _lib.is_channels_last(dy))):
                vec = 16 // dy.element_size()
                c = dy.shape[1]
                if c % vec == 0 and c // vec <= 256 and (y is not None or (act == 'linear' and clamp < 0)):
                    dyc = dy.contiguous(memory_format=torch.channels_last)
                    dx = torch.empty_like(dyc, memory_format=torch.channels_last)
                    dbf = torch.empty([c], dtype=torch.float32, device=dy.device)
                    n, _, h, w = dy.shape
                    with torch.cuda.device(dy.device), _lib.prof('bias_act_bwd_fused', 0.0, 3 * dy.numel() * dy.element_size()):
                        rc = _lib.lib().sgb_fused_epilogue_bwd(_lib.ptr(dyc), _lib.ptr(y), _lib.ptr(dx), None, None, None,
                                                               _lib.ptr(dbf), None, None, _lib.dtype_code(dy), n, c, h * w,
                                                               spec.cuda_idx, alpha, gain, clamp, _lib.stream_ptr(dy.device))
                    _lib.check(rc, 'fused_epilogue_bwd')
                    return dx, dbf.to(dy.dtype)
            if ctx.needs_input_grad[0] or (ctx.has_b and ctx.needs_input_grad[1]):
                dx = dy
                if not trivial:
                    dx = BiasActGrad.apply(dy, x, b, y)
            if ctx.has_b and ctx.needs_input_grad[1]:
                dxc, _ = _dense_like(dx)
                db = _SumToBias.apply(dxc)
            return dx, db

    class _SumToBias(torch.autograd.Function):      # differentiable so that d(db)/d(dy) exists under create_graph
        @staticmethod
        def forward(ctx, dx):
            ctx.shape = dx.shape
            return _sum_to_bias(dx, dim)

        @staticmethod
        def backward(ctx, g):
            view = [1] * len(ctx.shape)
            view[dim] = -1
            return g.reshape(view).expand(ctx.shape)

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            like = y if y is not None else x
            if like is not None:
                fmt = torch.channels_last if (like.dim() == 4 and _lib.is_channels_last(like)) else torch.contiguous_format
                dy = dy.contiguous(memory_format=fmt) if fmt == torch.channels_last else dy.contiguous()
            else:
                dy, fmt = _dense_like(dy)
            dx = _launch(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                like = dy
                d_dx_c = d_dx.contiguous(memory_format=torch.channels_last) if (like.dim() == 4 and _lib.is_channels_last(like)) else d_dx.contiguous()
                d_x = _launch(d_dx_c, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
                if b is not None and ctx.needs_input_grad[2]:
                    d_b = _sum_to_bias(d_x, dim)
            return d_dy, d_x, d_b, None

    _cache[key] = BiasAct
    return BiasAct


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain, -clamp, clamp); same arguments as the reference (bias_act.py:55-89).

    `impl` is accepted for signature compatibility ('ref' or 'cuda'); both run the CUDA kernels.
    Supports first and second order gradients."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    assert clamp is None or clamp >= 0
    _lib.require_cuda(x)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
        if b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError('b must have the same dtype and device as x')     # bias_act.cpp:36
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)
