"""modulated_conv2d: style modulation + convolution + weight demodulation (+ noise).

Same signature and semantics as the reference's `modulated_conv2d`
(`train_parts/generators.py:42-100`, identical copy in `stylegan2ada/training/networks.py:26-84`), built on the
libsgb200 kernels:

  * the modulation `x * styles` (reference :80) is not a separate pass: the convolution kernel multiplies the
    activations by `styles[n, c]` while it loads them (`in_scale`);
  * the demodulation coefficients are computed WITHOUT materialising the per-sample weights
    `w * styles` ([N,O,I,kh,kw], reference :70-74) through the identity
        dcoefs[n,o] = rsqrt( sum_i styles[n,i]^2 * (sum_k weight[o,i,k]^2) + 1e-8 )
    on [N,I] x [I,O] sized tensors;
  * `x * dcoefs + noise` (reference :83, fma.py) is one fused pass (`sgb_scale_nc`).

Every step is differentiable to any order (path-length regularisation differentiates the gradient wrt the
styles), and `conv2d_gradfix.no_weight_gradients()` is honoured by the convolution.
"""
import numpy as np
import torch

from .ops import conv2d_resample as _cr
from .ops import fma as _fma


def modulated_conv2d(
    x,                          # [batch_size, in_channels, in_height, in_width]
    weight,                     # [out_channels, in_channels, kernel_height, kernel_width]
    styles,                     # [batch_size, in_channels]
    noise           = None,     # optional noise added to the output, broadcastable to [N, 1, H, W]
    up              = 1,
    down            = 1,
    padding         = 0,
    resample_filter = None,
    demodulate      = True,
    flip_weight     = True,
    fused_modconv   = True,
    epilogue        = None,     # extension (default off): (bias, act, gain, clamp) of the bias_act that follows; fused into the
                                # convolution kernel together with demodulation and noise (ops/fused_conv.py)
):
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    assert weight.ndim == 4 and x.ndim == 4 and x.shape[1] == in_channels
    assert tuple(styles.shape) == (batch_size, in_channels)

    # fp16: pre-normalise so that neither the modulated activations nor the demodulation overflow (reference :63-65)
    if x.dtype == torch.float16 and demodulate:
        weight = weight * (1 / np.sqrt(in_channels * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)

    dcoefs = None
    if demodulate:
        wsq = weight.square().sum(dim=[2, 3])                                  # [O, I]
        dcoefs = (styles.square() @ wsq.t() + 1e-8).rsqrt()                    # [N, O]

    if not fused_modconv and epilogue is not None:
        b, act, gain, clamp = epilogue
        ep = _cr.Epilogue(b=b, act=act, gain=gain, clamp=clamp, dcoefs=dcoefs, noise=noise)
        return _cr.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                   flip_weight=flip_weight, in_scale=styles, epilogue=ep)
    if not fused_modconv:
        x = _cr.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                flip_weight=flip_weight, in_scale=styles)
        if noise is not None and (noise.ndim < 4 or noise.shape[0] != batch_size):
            noise = noise.expand(batch_size, 1, x.shape[2], x.shape[3])
        if demodulate:
            # the noise stays in its own (fp32) type: the kernel adds it in the accumulator type, and its gradient (the channel
            # sum of dy) is returned in fp32 instead of being rounded to fp16 first as `noise.to(x.dtype)` would make it
            return _fma.scale_nc(x, dcoefs, noise)
        if noise is not None:
            return x.add_(noise.to(x.dtype))
        return x

    # grouped convolution with one group per sample (reference :91-99); used in eval / G_ema mode
    w = weight.unsqueeze(0) * styles.reshape(batch_size, 1, -1, 1, 1)          # [N,O,I,kh,kw]
    if demodulate:
        w = w * dcoefs.reshape(batch_size, -1, 1, 1, 1)
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, in_channels, kh, kw)
    x = _cr.conv2d_resample(x=x, w=w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                            groups=batch_size, flip_weight=flip_weight)
    x = x.reshape(batch_size, -1, *x.shape[2:])
    if noise is not None:
        x = x.add_(noise)
    if epilogue is not None:
        from .ops import bias_act as _ba
        b, act, gain, clamp = epilogue
        x = _ba.bias_act(x, b, act=act, gain=gain, clamp=clamp)
    return x
