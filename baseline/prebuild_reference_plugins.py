"""Build the reference's own CUDA plugins (bias_act_plugin, upfirdn2d_plugin) from the snapshot baseline/_ref through the
reference's own custom_ops.get_plugin, into baseline/_ref/_torch_extensions (git-ignored, travels to the GPU box), so that
the "kernel to beat" runs on the box do not spend GPU minutes on a JIT build.  nvcc cross-compiles sm_100a without a GPU.

    python baseline/prebuild_reference_plugins.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prebuild(quiet=False):
    from benchmarks import ref_harness
    if ref_harness.reference_root() is None:
        return None
    ref_harness.import_reference('reference')
    from stylegan2ada.torch_utils import custom_ops
    custom_ops.verbosity = 'none' if quiet else 'brief'
    st = ref_harness.plugin_status()
    if not quiet:
        print('reference CUDA plugins built:', st)
    return st


if __name__ == '__main__':
    st = prebuild()
    raise SystemExit(0 if st and all(st.values()) else 1)
