"""Put the sgb200 ops in front of UNCHANGED reference callers.

The reference imports its ops by dotted name (`from stylegan2ada.torch_utils.ops import bias_act`,
train_parts/generators.py:19-22, discriminators.py:19-22, regularizations.py:4, losses_base.py:13) and defines
`modulated_conv2d` as a module global of `train_parts.generators` / `stylegan2ada.training.networks`
(generators.py:42, looked up at call time from SynthesisLayer.forward, :323).  `install()`:

  1. registers our modules in `sys.modules` as `stylegan2ada.torch_utils.ops.{bias_act,upfirdn2d,conv2d_resample,
     conv2d_gradfix,fma,grid_sample_gradfix}` (and as attributes of the reference's `ops` package when it is
     importable), so every later `from stylegan2ada.torch_utils.ops import X` resolves to them;
  2. installs a `sys.meta_path` post-import hook that rebinds `modulated_conv2d` in `train_parts.generators` and
     `stylegan2ada.training.networks` right after those modules finish executing (their own `def modulated_conv2d`
     runs after the op imports), and rebinds it immediately in modules that are already imported.

Ranks started with `spawn` (starter.py:25-30) re-import everything: call `install()` at the top of the rank entry,
or put `style-big-gan_b200/overlay` on PYTHONPATH ahead of the reference checkout -- the overlay's
`sitecustomize`-free package shim calls `install()` on first import of `stylegan2ada.torch_utils.ops`.
"""
import importlib
import importlib.abc
import importlib.util
import sys

OPS = ('bias_act', 'upfirdn2d', 'conv2d_resample', 'conv2d_gradfix', 'fma', 'grid_sample_gradfix')
REBIND = ('train_parts.generators', 'stylegan2ada.training.networks')
_PREFIX = 'stylegan2ada.torch_utils.ops'
_installed = False


def _rebind(module):
    from .modconv import modulated_conv2d
    if hasattr(module, 'modulated_conv2d'):
        module._reference_modulated_conv2d = module.modulated_conv2d
        module.modulated_conv2d = modulated_conv2d


class _PostImportHook(importlib.abc.MetaPathFinder):
    """Wraps the real loader of the modules in REBIND so that `_rebind` runs after exec_module."""

    def find_spec(self, fullname, path, target=None):
        if fullname not in REBIND:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, 'find_spec'):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _WrappedLoader(spec.loader)
                return spec
        return None


class _WrappedLoader(importlib.abc.Loader):
    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        _rebind(module)

    def __getattr__(self, name):
        return getattr(self.inner, name)


def install():
    """Idempotent.  Returns the dict {dotted name: module} that was registered."""
    global _installed
    from . import ops as our_ops
    registered = {}
    try:                                   # reference checkout on sys.path: keep its package objects, swap the leaves
        pkg = importlib.import_module(_PREFIX)
    except Exception:
        pkg = None
    for name in OPS:
        mod = getattr(our_ops, name)
        sys.modules[f'{_PREFIX}.{name}'] = mod
        registered[f'{_PREFIX}.{name}'] = mod
        if pkg is not None:
            setattr(pkg, name, mod)
    if not _installed:
        sys.meta_path.insert(0, _PostImportHook())
        _installed = True
    for name in REBIND:
        if name in sys.modules:
            _rebind(sys.modules[name])
    return registered
