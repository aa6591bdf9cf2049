"""upfirdn2d: pad, zero-insert upsample, FIR filter, decimate -- on the sm_100a kernels of libsgb200.

Public interface identical to the reference's `stylegan2ada/torch_utils/ops/upfirdn2d.py`:
`setup_filter` (:72), `upfirdn2d` (:120), `filter2d` (:272), `upsample2d` (:308), `downsample2d` (:347) and the
private helpers `_parse_padding` / `_get_filter_size` / `_parse_scaling` that `conv2d_resample` imports.
Gradients of arbitrary order: the backward of the op is the op itself with up/down swapped and the filter
flipped (reference :246-264).
"""
import ctypes
import math
import os
import weakref

import numpy as np
import torch

from .. import _lib


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple))
    assert all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    padx0, padx1, pady0, pady1 = padding
    return padx0, padx1, pady0, pady1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw = int(f.shape[-1])
    fh = int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Build the fp32 FIR filter tensor the way the reference does (upfirdn2d.py:72-116): taps given as a
    1-D list with fewer than 8 entries become their 2-D outer product, longer ones stay separable."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


separable_kernel = os.environ.get('SGB_FIR_SEP', '1') != '0'     # A/B switch: the separable strip kernel for plain FIR passes
_sep_cache = {}     # (data_ptr, shape, strides, version) -> (the tensor, (row, col) or None)
stats = dict(sep=0, general=0, unseen_in_capture=0)     # launches by kernel (tests)


def _separable_taps(f2d, flip, gain):
    """(fx, fy) as 4-tap host lists in application order if the (at most 4 x 4) filter is an outer product, else None.
    The factorisation needs the filter values on the host: one device->host copy per filter BUFFER (keyed by address and
    version counter -- autograd hands the backward a new Python object for the saved filter, so identity of the object is
    not enough; the entry keeps the tensor alive, so the address cannot be recycled), never during CUDA-graph capture (an
    unseen filter then takes the general kernel)."""
    fh, fw = f2d.shape
    if fh > 4 or fw > 4:
        return None
    key = (f2d.data_ptr(), fh, fw, f2d.stride(0), f2d.stride(1), f2d._version)
    hit = _sep_cache.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            stats['unseen_in_capture'] += 1
            return None
        a = f2d.detach().to('cpu', torch.float64).numpy()
        fac = None
        r0, c0 = np.unravel_index(np.argmax(np.abs(a)), a.shape)
        if a[r0, c0] != 0:
            row, col = a[r0, :].copy(), a[:, c0] / a[r0, c0]
            if np.abs(np.outer(col, row) - a).max() <= 1e-7 * np.abs(a[r0, c0]):
                fac = (row, col)
        if len(_sep_cache) > 256:
            _sep_cache.clear()
        hit = (f2d.detach(), f2d._version, fac)
        _sep_cache[key] = hit
    if hit[2] is None:
        return None
    row, col = hit[2]
    # application order (csrc/upfirdn2d.cu): tap t multiplies the filter entry t if flip_filter else size-1-t; unused taps are zero
    fx = [float(row[t] if flip else row[fw - 1 - t]) * float(gain) if t < fw else 0.0 for t in range(4)]
    fy = [float(col[t] if flip else col[fh - 1 - t]) if t < fh else 0.0 for t in range(4)]
    return fx, fy


def _run(x, f2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """One sgb_upfirdn2d launch with a 2-D fp32 filter (any strides)."""
    n, c, ih, iw = x.shape
    fh, fw = f2d.shape
    ow = (iw * upx + padx0 + padx1 - fw + downx) // downx
    oh = (ih * upy + pady0 + pady1 - fh + downy) // downy
    if ow < 1 or oh < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')          # upfirdn2d.cpp:33
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=_lib.out_format(x))
    if y.numel() == 0:
        return y
    vec = 16 // x.element_size()
    if (separable_kernel and upx == upy == downx == downy == 1 and x.dtype in (torch.float32, torch.float16, torch.bfloat16)
            and c % vec == 0 and _lib.is_channels_last(x) and _lib.is_channels_last(y) and x.data_ptr() % 16 == 0
            and all(st % vec == 0 for i, st in enumerate(x.stride()) if i != 1)):
        taps = _separable_taps(f2d, flip, gain)
        if taps is not None:
            fx = (ctypes.c_float * 4)(*taps[0])
            fy = (ctypes.c_float * 4)(*taps[1])
            with torch.cuda.device(x.device), _lib.prof('upfirdn2d', 0.0, (x.numel() + y.numel()) * x.element_size()):
                rc = _lib.lib().sgb_upfirdn2d_sep(_lib.ptr(x), fx, fy, _lib.ptr(y), _lib.dtype_code(x), n, c, ih, iw, _lib.strides4(x),
                                                  oh, ow, _lib.strides4(y), padx0, pady0, _lib.stream_ptr(x.device))
            _lib.check(rc, 'upfirdn2d_sep')
            stats['sep'] += 1
            return y
    with torch.cuda.device(x.device), _lib.prof('upfirdn2d', 0.0, (x.numel() + y.numel()) * x.element_size()):
        rc = _lib.lib().sgb_upfirdn2d(_lib.ptr(x), _lib.ptr(f2d), _lib.ptr(y), _lib.dtype_code(x),
                                      n, c, ih, iw, _lib.strides4(x), oh, ow, _lib.strides4(y),
                                      fh, fw, f2d.stride(0), f2d.stride(1),
                                      upx, upy, downx, downy, padx0, pady0, int(bool(flip)), float(gain),
                                      _lib.stream_ptr(x.device))
    _lib.check(rc, 'upfirdn2d')
    stats['general'] += 1
    return y


_cache = dict()


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _cache:
        return _cache[key]

    class Upfirdn2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
            if f.dtype != torch.float32:
                raise RuntimeError('f must be float32')                       # upfirdn2d.cpp:21
            if f.device != x.device:
                raise RuntimeError('f must reside on the same device as x')   # upfirdn2d.cpp:20
            if f.ndim == 2:
                y = _run(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:   # separable: along x, then along y (reference :236-237)
                g = math.sqrt(gain)
                y = _run(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, g)
                y = _run(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, g)
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [
                fw - padx0 - 1,
                iw * upx - ow * downx + padx0 - upx + 1,
                fh - pady0 - 1,
                ih * upy - oh * downy + pady0 - upy + 1,
            ]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _cache[key] = Upfirdn2d
    return Upfirdn2d


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, upsample, filter and downsample a batch of 2-D images (reference upfirdn2d.py:120-164).
    `impl` is accepted for compatibility; both values run the CUDA kernels."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    _lib.require_cuda(x)
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Same-size FIR filtering (reference :272-304)."""
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample by an integer factor with the FIR filter (reference :308-343)."""
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Downsample by an integer factor with the FIR filter (reference :347-382)."""
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
