"""conv2d_resample: 2-D convolution with optional FIR up / down sampling.

Interface of the reference's `stylegan2ada/torch_utils/ops/conv2d_resample.py:59`
(`conv2d_resample(x, w, f, up, down, padding, groups, flip_weight, flip_filter)`); same padding algebra
(:94-104) and the same decomposition into the six cases (:106-154), executed by libsgb200's convolution and
upfirdn2d kernels.  `flip_weight=False` is passed to the kernel as a flag instead of materialising
`w.flip([2, 3])` (:35-36).  Extension (keyword-only): `in_scale=[N,Ci]` fuses the style modulation
`x * styles` (generators.py:80) into the convolution's operand load.
"""
import torch

from . import conv2d_gradfix
from . import upfirdn2d
from . import fma as _fma
from .upfirdn2d import _parse_padding
from .upfirdn2d import _get_filter_size


def _get_weight_shape(w):
    return [int(sz) for sz in w.shape]


class Epilogue:
    """What the layer runs after the convolution (extension, default off): `y * dcoefs[n,c] + noise`, then
    bias_act(b, act, gain, clamp).  When the convolution is the LAST stage of the resampling decomposition it is fused
    into the convolution kernel (ops/fused_conv.py); otherwise it is applied with the stand-alone ops."""
    def __init__(self, b=None, act='linear', alpha=None, gain=None, clamp=None, dcoefs=None, noise=None):
        self.b, self.act, self.alpha, self.gain, self.clamp, self.dcoefs, self.noise = b, act, alpha, gain, clamp, dcoefs, noise

    def apply(self, y):
        from . import bias_act
        if self.dcoefs is not None or self.noise is not None:
            from . import fused_conv           # demodulation + noise + bias_act as one pass (and a one-pass backward)
            return fused_conv.scale_bias_act(y, self.b, self.dcoefs, self.noise, act=self.act, alpha=self.alpha, gain=self.gain,
                                             clamp=self.clamp)
        noise = self.noise
        if noise is not None:
            noise = noise.to(y.dtype)
            if noise.ndim < 4 or noise.shape[0] != y.shape[0]:
                noise = noise.expand(y.shape[0], 1, y.shape[2], y.shape[3])
        if self.dcoefs is not None:
            y = _fma.scale_nc(y, self.dcoefs, noise)
        elif noise is not None:
            y = y.add_(noise)
        return bias_act.bias_act(y, self.b, act=self.act, alpha=self.alpha, gain=self.gain, clamp=self.clamp)


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True, in_scale=None, epilogue=None):
    """conv2d() is a correlation: flip_weight=True means "use w as is" (reference :29-54)."""
    if epilogue is not None:
        if not transpose and groups == 1:
            from . import fused_conv
            e = epilogue
            return fused_conv.conv2d_bias_act(x, w, e.b, stride=stride, padding=padding, flip_weight=flip_weight, styles=in_scale,
                                              dcoefs=e.dcoefs, noise=e.noise, act=e.act, alpha=e.alpha, gain=e.gain, clamp=e.clamp)
        return epilogue.apply(_conv2d_wrapper(x, w, stride, padding, groups, transpose, flip_weight, in_scale))
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups, flip_weight=(not flip_weight), in_scale=in_scale)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False, *, in_scale=None,
                    epilogue=None):
    assert isinstance(x, torch.Tensor) and (x.ndim == 4)
    assert isinstance(w, torch.Tensor) and (w.ndim == 4) and (w.dtype == x.dtype)
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and (up >= 1)
    assert isinstance(down, int) and (down >= 1)
    assert isinstance(groups, int) and (groups >= 1)
    out_channels, in_channels_per_group, kh, kw = _get_weight_shape(w)
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)

    # padding seen by the resampling filter (reference :94-104)
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
    conv = dict(groups=groups, flip_weight=flip_weight, epilogue=epilogue)      # cases whose last stage is the convolution
    tail = (lambda y: y) if epilogue is None else epilogue.apply                 # cases that end with a FIR pass

    # 1x1 conv + downsampling: filter/decimate first, then convolve
    if kw == 1 and kh == 1 and (down > 1 and up == 1):
        if in_scale is not None:
            x = _fma.scale_nc(x, in_scale)
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, **conv)

    # 1x1 conv + upsampling: convolve first, then upsample
    if kw == 1 and kh == 1 and (up > 1 and down == 1):
        x = _conv2d_wrapper(x=x, w=w, in_scale=in_scale, groups=groups, flip_weight=flip_weight)
        return tail(upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter))

    # downsampling only: FIR at full resolution, then strided conv
    if down > 1 and up == 1:
        if in_scale is not None:
            x = _fma.scale_nc(x, in_scale)
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, **conv)

    # upsampling (optionally followed by downsampling): transposed strided conv, then FIR
    if up > 1:
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_channels_per_group, kh, kw)
            w = w.transpose(1, 2)
            w = w.reshape(groups * in_channels_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x=x, w=w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True,
                            flip_weight=(not flip_weight), in_scale=in_scale)
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return tail(x)

    # no resampling and a padding the conv kernel takes directly
    if up == 1 and down == 1:
        if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
            return _conv2d_wrapper(x=x, w=w, padding=[py0, px0], in_scale=in_scale, **conv)

    # generic: explicit pad / upsample, conv, downsample
    if in_scale is not None:
        x = _fma.scale_nc(x, in_scale)
    x = upfirdn2d.upfirdn2d(x=x, f=(f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1:
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return tail(upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter))
    return _conv2d_wrapper(x=x, w=w, **conv)
