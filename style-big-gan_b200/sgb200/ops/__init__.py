"""Drop-in replacements for `stylegan2ada.torch_utils.ops.*` backed by libsgb200 (sm_100a CUDA)."""
from . import bias_act, upfirdn2d, conv2d_gradfix, conv2d_resample, fma, grid_sample_gradfix  # noqa: F401
