"""fma(a, b, c) = a * b + c and the activation-sized modulation passes, on libsgb200 kernels.

`fma` has the interface of the reference's `stylegan2ada/torch_utils/ops/fma.py:15` (used by
`modulated_conv2d` for `x * dcoefs + noise`, generators.py:83).  The shape the hot path uses --
a [N,C,H,W], b [N,C,1,1], c [N,1,H,W] -- runs on `sgb_scale_nc`; its gradients are the reductions the
reference performs with `_unbroadcast` (fma.py:49-58), here `sgb_mul_sum_hw` and `sgb_sum_c`.  Every op is an
autograd.Function whose backward is built from the others, so gradients of any order exist.
"""
import torch

from .. import _lib


def _raw_scale(x, s, t):
    y = torch.empty_like(x, memory_format=_lib.out_format(x))
    if y.numel() == 0:
        return y
    n, c, h, w = x.shape
    acc = _lib.acc_dtype(x.dtype)
    s = s.detach().reshape(n, c).to(acc).contiguous()
    if t is not None:
        t = t.detach().reshape(n, h, w).to(acc).contiguous()
    with torch.cuda.device(x.device), _lib.prof('scale_nc', 0.0, 2 * x.numel() * x.element_size()):
        rc = _lib.lib().sgb_scale_nc(_lib.ptr(x), _lib.ptr(s), _lib.ptr(t), _lib.ptr(y), _lib.dtype_code(x), n, c, h, w,
                                     _lib.strides4(x), _lib.strides4(y), _lib.stream_ptr(x.device))
    _lib.check(rc, 'scale_nc')
    return y


def _raw_mul_sum_hw(a, b):
    n, c, h, w = a.shape
    out = torch.empty([n, c], dtype=_lib.acc_dtype(a.dtype), device=a.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(a.device), _lib.prof('mul_sum_hw', 0.0, 2 * a.numel() * a.element_size()):
        rc = _lib.lib().sgb_mul_sum_hw(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.dtype_code(a), n, c, h, w,
                                       _lib.strides4(a), _lib.strides4(b), _lib.stream_ptr(a.device))
    _lib.check(rc, 'mul_sum_hw')
    return out


def _raw_sum_c(a):
    n, c, h, w = a.shape
    out = torch.empty([n, 1, h, w], dtype=_lib.acc_dtype(a.dtype), device=a.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(a.device), _lib.prof('sum_c', 0.0, a.numel() * a.element_size()):
        rc = _lib.lib().sgb_sum_c(_lib.ptr(a), _lib.ptr(out), _lib.dtype_code(a), n, c, h, w, _lib.strides4(a),
                                  _lib.stream_ptr(a.device))
    _lib.check(rc, 'sum_c')
    return out


class _ScaleNC(torch.autograd.Function):
    """y = x * s[n,c] (+ t[n,1,h,w])."""
    @staticmethod
    def forward(ctx, x, s, t):
        ctx.save_for_backward(x, s)
        ctx.has_t = t is not None
        ctx.t_dtype = t.dtype if t is not None else None
        ctx.s_shape = s.shape
        return _raw_scale(x, s, t)

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        dx = ds = dt = None
        if ctx.needs_input_grad[0]:
            dx = _ScaleNC.apply(dy, s, None)
        if ctx.needs_input_grad[1]:
            ds = _MulSumHW.apply(dy, x).to(s.dtype).reshape(ctx.s_shape)
        if ctx.has_t and ctx.needs_input_grad[2]:
            dt = _SumC.apply(dy).to(ctx.t_dtype)
        return dx, ds, dt


class _MulSumHW(torch.autograd.Function):
    """out[n,c] = sum_hw a*b (accumulator dtype)."""
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _raw_mul_sum_hw(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _ScaleNC.apply(b, g, None)
        if ctx.needs_input_grad[1]:
            db = _ScaleNC.apply(a, g, None)
        return da, db


class _SumC(torch.autograd.Function):
    """out[n,1,h,w] = sum_c a (accumulator dtype)."""
    @staticmethod
    def forward(ctx, a):
        ctx.shape = a.shape
        ctx.dtype = a.dtype
        return _raw_sum_c(a)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).expand(ctx.shape)


def scale_nc(x, s, t=None):
    """x[N,C,H,W] * s[N,C] (+ t[N,1,H,W] or [N,H,W]); differentiable to any order in x, s and t."""
    _lib.require_cuda(x)
    if t is not None and t.ndim == 3:
        t = t.unsqueeze(1)
    return _ScaleNC.apply(x, s, t)


def mul_sum_hw(a, b):
    return _MulSumHW.apply(a, b)


def sum_c(a):
    return _SumC.apply(a)


def _unbroadcast(x, shape):
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert x.shape == shape
    return x


def fma(a, b, c):
    """a * b + c.  The modulated-conv shape runs on the fused CUDA kernel; any other broadcast pattern is
    evaluated with `torch.addcmul` (same values, differentiable) since it is not on the hot path."""
    _lib.require_cuda(a, 'a')
    if (a.ndim == 4 and b.ndim == 4 and c.ndim == 4 and b.shape == (a.shape[0], a.shape[1], 1, 1)
            and c.shape == (a.shape[0], 1, a.shape[2], a.shape[3])):
        return _ScaleNC.apply(a, b.reshape(a.shape[0], a.shape[1]), c)
    return torch.addcmul(c, a, b)
