"""ctypes binding of libsgb200.so (C ABI declared in include/sgb200.h).

There is NO fallback: if the library is missing, cannot be loaded, or a call returns non-zero, a
RuntimeError is raised.  PyTorch owns every buffer and the stream; this module only passes raw pointers.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libsgb200.so')

SGB_F32, SGB_F16, SGB_BF16, SGB_F64 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: SGB_F32, torch.float16: SGB_F16, torch.bfloat16: SGB_BF16, torch.float64: SGB_F64}

_c = ctypes
_i64 = _c.c_int64
_vp = _c.c_void_p
_int = _c.c_int
_flt = _c.c_float
_I64x4 = _i64 * 4


class ConvDesc(_c.Structure):
    """Mirror of sgb_conv_desc (include/sgb200.h)."""
    _fields_ = [
        ('dtype', _c.c_int32), ('transposed', _c.c_int32),
        ('n', _c.c_int32), ('ci', _c.c_int32), ('co', _c.c_int32),
        ('in_h', _c.c_int32), ('in_w', _c.c_int32), ('out_h', _c.c_int32), ('out_w', _c.c_int32),
        ('kh', _c.c_int32), ('kw', _c.c_int32), ('stride', _c.c_int32),
        ('pad_y', _c.c_int32), ('pad_x', _c.c_int32), ('groups', _c.c_int32), ('flip', _c.c_int32),
        ('x_strides', _I64x4), ('y_strides', _I64x4),
        ('in_scale', _vp), ('out_scale', _vp), ('noise', _vp), ('bias', _vp),
        ('act', _c.c_int32), ('alpha', _flt), ('gain', _flt), ('clamp', _flt),
        ('strict_fp32', _c.c_int32), ('force_simt', _c.c_int32), ('halo_gt', _c.c_int32), ('workspace', _vp), ('workspace_bytes', _i64),
    ]


# name -> (restype, argtypes); must list every symbol include/sgb200.h declares
SIGNATURES = {
    'sgb_last_error': (_c.c_char_p, []),
    'sgb_abi_version': (_int, []),
    'sgb_launch_count': (_i64, []),
    'sgb_bias_act': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _flt, _flt, _flt, _i64, _i64, _i64, _vp]),
    'sgb_sum_to_channel': (_int, [_vp, _vp, _int, _i64, _i64, _i64, _vp]),
    'sgb_upfirdn2d': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _c.POINTER(_i64), _int, _int, _c.POINTER(_i64),
                             _int, _int, _i64, _i64, _int, _int, _int, _int, _int, _int, _int, _flt, _vp]),
    'sgb_upfirdn2d_sep': (_int, [_vp, _c.POINTER(_flt), _c.POINTER(_flt), _vp, _int, _int, _int, _int, _int, _c.POINTER(_i64), _int, _int,
                          _c.POINTER(_i64), _int, _int, _vp]),
    'sgb_conv2d_forward': (_int, [_c.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    'sgb_conv2d_wgrad': (_int, [_c.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    'sgb_conv2d_uses_tensor_cores': (_int, [_c.POINTER(ConvDesc)]),
    'sgb_conv2d_wgrad_uses_tensor_cores': (_int, [_c.POINTER(ConvDesc)]),
    'sgb_conv2d_workspace_bytes': (_i64, [_c.POINTER(ConvDesc)]),
    'sgb_scale_nc': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _c.POINTER(_i64), _c.POINTER(_i64), _vp]),
    'sgb_mul_sum_hw': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _c.POINTER(_i64), _c.POINTER(_i64), _vp]),
    'sgb_sum_c': (_int, [_vp, _vp, _int, _int, _int, _int, _int, _c.POINTER(_i64), _vp]),
    'sgb_scale_bias_act': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _flt, _flt, _flt, _vp]),
    'sgb_nan_to_num_multi': (_int, [_c.POINTER(_c.c_void_p), _c.POINTER(_i64), _int, _flt, _flt, _flt, _vp]),
    'sgb_mod_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    'sgb_fused_epilogue_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _flt, _flt, _flt, _vp]),
}

_lib = None


def lib():
    """The loaded library; raises RuntimeError (never falls back) when it is absent or incomplete."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'sgb200: CUDA library not built: {LIB_PATH} is missing '
                               f'(run `make -C style-big-gan_b200/csrc` or __graft_entry__.build()); there is no fallback')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                raise RuntimeError(f'sgb200: {LIB_PATH} does not export {name}')
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().sgb_last_error()
        raise RuntimeError(f'sgb200 {what}: {msg.decode() if msg else "error %d" % rc}')


def dtype_code(t):
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f'sgb200: unsupported dtype {t.dtype}')


def acc_dtype(dtype):
    """dtype of scales / reduction outputs for tensors of `dtype` (fp32, fp64 for fp64)."""
    return torch.float64 if dtype == torch.float64 else torch.float32


def require_cuda(t, name='x'):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'sgb200: {name} must be a CUDA tensor; this package has no CPU implementation')


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def strides4(t):
    return _I64x4(*t.stride())


def launch_count():
    return int(lib().sgb_launch_count())


def is_channels_last(t):
    """True when a 4-D tensor's channel dimension is the contiguous one (and that is not trivially so)."""
    return t.dim() == 4 and t.shape[1] > 1 and t.stride(1) == 1 and t.is_contiguous(memory_format=torch.channels_last)


def out_format(t):
    return torch.channels_last if is_channels_last(t) else torch.contiguous_format


# ---- optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream ----------
class _Profile:
    def __init__(self):
        self.records = []          # (kind, start_event, end_event, flops, bytes, tag)

    def summary(self, by_tag=False):
        """kind (or "kind | tag" with by_tag) -> dict(launches, ms, flops, bytes); call after torch.cuda.synchronize()."""
        out = {}
        for kind, e0, e1, fl, by, tag in self.records:
            d = out.setdefault(f'{kind} | {tag}' if (by_tag and tag) else kind, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d['launches'] += 1
            d['ms'] += e0.elapsed_time(e1)
            d['flops'] += fl
            d['bytes'] += by
        return out


PROFILE = None


def profile_start():
    global PROFILE
    PROFILE = _Profile()
    return PROFILE


def profile_stop():
    global PROFILE
    p, PROFILE = PROFILE, None
    return p


class prof:
    """`with prof(kind, flops, bytes):` around ONE kernel launch; free when profiling is off."""
    __slots__ = ('kind', 'flops', 'bytes', 'e0', 'tag')

    def __init__(self, kind, flops=0.0, nbytes=0.0, tag=None):
        self.kind, self.flops, self.bytes, self.tag = kind, flops, nbytes, tag

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.records.append((self.kind, self.e0, e1, float(self.flops), float(self.bytes), self.tag))
        return False
