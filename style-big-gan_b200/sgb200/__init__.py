"""sgb200: B200-native (sm_100a) StyleGAN2-ADA op hot path behind the reference's own op API.

    from sgb200.ops import bias_act, upfirdn2d, conv2d_resample, conv2d_gradfix, fma
    from sgb200.modconv import modulated_conv2d

`sgb200.install()` puts these modules in front of unchanged reference callers
(`stylegan2ada.torch_utils.ops.*`, `modulated_conv2d` in `train_parts.generators`); see INTEGRATION.md.
"""
from . import _lib            # noqa: F401
from . import ops             # noqa: F401
from .modconv import modulated_conv2d   # noqa: F401

__version__ = '0.1.0'


def install(*a, **k):
    # the implementation lives in `_install.py`: a submodule named `install` would replace this function as the package
    # attribute `sgb200.install` the first time it is imported, and the second `sgb200.install()` would fail
    from ._install import install as _impl
    return _impl(*a, **k)
