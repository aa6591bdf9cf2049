"""CPU oracle for the StyleGAN2-ADA op hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain PyTorch CPU arithmetic, of what the reference's
``impl='ref'`` path computes.  It is the *checker* for the CUDA kernels: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``style-big-gan_b200/`` imports it and the product
path raises when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the real
modules from ``/root/reference`` on CPU and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.

Third-party arithmetic: ``torch.nn.functional.conv2d`` / ``conv_transpose2d`` (the reference
pins torch==1.7.1 in requirements.txt:1; this image has torch 2.11).  All citations are
relative to ``/root/reference``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# bias_act  (stylegan2ada/torch_utils/ops/bias_act.py:23-33 table, :93-123 ref impl)

# name -> (function, default alpha, default gain)
_ACTS = {
    'linear':   (lambda v, a: v,                         0.0, 1.0),
    'relu':     (lambda v, a: torch.relu(v),             0.0, math.sqrt(2.0)),
    'lrelu':    (lambda v, a: F.leaky_relu(v, a),        0.2, math.sqrt(2.0)),
    'tanh':     (lambda v, a: torch.tanh(v),             0.0, 1.0),
    'sigmoid':  (lambda v, a: torch.sigmoid(v),          0.0, 1.0),
    'elu':      (lambda v, a: F.elu(v),                  0.0, 1.0),
    'selu':     (lambda v, a: F.selu(v),                 0.0, 1.0),
    'softplus': (lambda v, a: F.softplus(v),             0.0, 1.0),
    'swish':    (lambda v, a: torch.sigmoid(v) * v,      0.0, math.sqrt(2.0)),
}


def act_defaults(act):
    """(def_alpha, def_gain) as in bias_act.py:23-33."""
    _, a, g = _ACTS[act]
    return a, g


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """y = clamp(act(x + b) * gain).  Follows bias_act.py:93-123 step by step."""
    fn, def_alpha, def_gain = _ACTS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    if b is not None:
        assert b.ndim == 1 and b.shape[0] == x.shape[dim]
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    y = fn(x, alpha)
    if gain != 1:
        y = y * gain
    if clamp is not None and clamp >= 0:
        y = y.clamp(-float(clamp), float(clamp))
    return y


# ----------------------------------------------------------------------------------------
# upfirdn2d  (stylegan2ada/torch_utils/ops/upfirdn2d.py)

def _pair(v):
    if isinstance(v, int):
        return v, v
    a, b = v
    return int(a), int(b)


def parse_padding(padding):
    """upfirdn2d.py:46-55 -> (padx0, padx1, pady0, pady1)."""
    if isinstance(padding, int):
        padding = [padding, padding]
    padding = list(padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    return tuple(int(p) for p in padding)


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:72-116: 1-D taps with < 8 entries become an outer product."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """zero-insert by `up`, pad/crop, FIR, keep every `down`-th sample (upfirdn2d.py:168-208)."""
    n, c, h, w = x.shape
    ux, uy = _pair(up)
    dx, dy = _pair(down)
    px0, px1, py0, py1 = parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    # zero insertion: sample (i, j) lands at (i*uy, j*ux)
    z = x.new_zeros([n, c, h * uy, w * ux])
    z[:, :, ::uy, ::ux] = x
    # positive padding adds zeros, negative padding crops
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    taps = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:           # true convolution = correlation with the flipped taps
        taps = taps.flip(list(range(taps.ndim)))
    if taps.ndim == 2:
        k = taps[None, None].repeat(c, 1, 1, 1)
        z = F.conv2d(z, k, groups=c)
    else:                         # separable: along x, then along y
        kx = taps[None, None, None, :].repeat(c, 1, 1, 1)
        ky = taps[None, None, :, None].repeat(c, 1, 1, 1)
        z = F.conv2d(z, kx, groups=c)
        z = F.conv2d(z, ky, groups=c)
    return z[:, :, ::dy, ::dx]


def _fsize(f):
    if f is None:
        return 1, 1
    return int(f.shape[-1]), int(f.shape[0])


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:272-304."""
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = _fsize(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:308-343."""
    ux, uy = _pair(up)
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:347-382."""
    dx, dy = _pair(down)
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# ----------------------------------------------------------------------------------------
# conv2d_resample  (stylegan2ada/torch_utils/ops/conv2d_resample.py:29-154)

def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:29-54 minus the cuDNN 8.0.5 workaround (a pure layout detour)."""
    if not flip_weight:
        w = w.flip([2, 3])
    if transpose:
        return F.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
    return F.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:58-154: same branch order, same padding algebra."""
    co, cig, kh, kw = (int(s) for s in w.shape)
    fw, fh = _fsize(f)
    px0, px1, py0, py1 = parse_padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2

    if kw == 1 and kh == 1 and down > 1 and up == 1:          # :107-110
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:          # :113-116
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                  # :119-122
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                                # :125-142
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, co // groups, cig, kh, kw).transpose(1, 2)
            wt = wt.reshape(groups * cig, co // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, wt, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # :145-147
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    # generic fallback :150-154
    x = upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


# ----------------------------------------------------------------------------------------
# fma  (stylegan2ada/torch_utils/ops/fma.py:15-23)

def fma(a, b, c):
    return torch.addcmul(c, a, b)


# ----------------------------------------------------------------------------------------
# modulated_conv2d  (train_parts/generators.py:42-100 == stylegan2ada/training/networks.py:26-84)

def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    n = x.shape[0]
    co, ci, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:               # :63-65 fp16 pre-normalisation
        weight = weight * (1 / np.sqrt(ci * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    wmod = None
    dco = None
    if demodulate or fused_modconv:                           # :70-72
        wmod = weight.unsqueeze(0) * styles.reshape(n, 1, ci, 1, 1)
    if demodulate:                                            # :73-74
        dco = (wmod.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:                          # :75-76
        wmod = wmod * dco.reshape(n, co, 1, 1, 1)

    if not fused_modconv:                                     # :79-88 scale activations before / after
        x = x * styles.to(x.dtype).reshape(n, ci, 1, 1)
        x = conv2d_resample(x, weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                            flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma(x, dco.to(x.dtype).reshape(n, co, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dco.to(x.dtype).reshape(n, co, 1, 1)
        elif noise is not None:
            x = x + noise.to(x.dtype)
        return x

    # :91-99 grouped convolution, one group per sample
    x = x.reshape(1, n * ci, *x.shape[2:])
    wg = wmod.reshape(n * co, ci, kh, kw)
    x = conv2d_resample(x, wg.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, groups=n,
                        flip_weight=flip_weight)
    x = x.reshape(n, co, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x
