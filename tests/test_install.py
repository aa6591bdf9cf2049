"""CPU: `sgb200.install()` / the path overlay put our ops in front of the UNCHANGED reference callers.
Needs the reference checkout (authoring container only); skipped on the GPU box."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')

PRELUDE = textwrap.dedent('''
    import sys, types, dataclasses, collections, collections.abc
    om = types.ModuleType('omegaconf'); om.MISSING = '???'; om.OmegaConf = type('OmegaConf', (), {})
    lc = types.ModuleType('omegaconf.listconfig'); lc.ListConfig = list; om.listconfig = lc
    sys.modules['omegaconf'] = om; sys.modules['omegaconf.listconfig'] = lc
    _orig = dataclasses.make_dataclass
    dataclasses.make_dataclass = lambda *a, **k: _orig(*a, **{**k, 'eq': k.get('eq', False)})
    collections.MutableMapping = collections.abc.MutableMapping
''')


def _run(code, pythonpath):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath))
    r = subprocess.run([sys.executable, '-c', PRELUDE + textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_explicit_install_rebinds_reference_callers():
    out = _run('''
        import sgb200
        sgb200.install()
        import train_parts.generators as G, train_parts.discriminators as D, train_parts.regularizations as Rg
        import sgb200.ops as ops
        from sgb200.modconv import modulated_conv2d
        assert G.bias_act is ops.bias_act and G.upfirdn2d is ops.upfirdn2d and G.conv2d_resample is ops.conv2d_resample
        assert D.bias_act is ops.bias_act and D.conv2d_resample is ops.conv2d_resample
        assert Rg.conv2d_gradfix is ops.conv2d_gradfix
        assert G.modulated_conv2d is modulated_conv2d, 'post-import hook did not rebind modulated_conv2d'
        assert G._reference_modulated_conv2d is not modulated_conv2d
        import stylegan2ada.training.networks as Nw
        assert Nw.modulated_conv2d is modulated_conv2d
        # the unchanged caller now reaches our op: CPU tensors must be refused (no fallback)
        import torch
        layer = G.Conv2dLayer(4, 4, 3)
        try:
            layer(torch.zeros(1, 4, 8, 8))
        except RuntimeError as e:
            assert 'no CPU implementation' in str(e)
        else:
            raise SystemExit('reference Conv2dLayer did not reach sgb200')
        print('OK')
    ''', [os.path.join(ROOT, 'style-big-gan_b200'), REF])
    assert 'OK' in out


def test_path_overlay():
    out = _run('''
        import stylegan2ada.torch_utils.ops as ops_pkg            # overlay package -> installs sgb200
        from stylegan2ada.torch_utils import misc, persistence     # still the reference's own modules
        assert misc.__file__.startswith('/root/reference'), misc.__file__
        import sgb200.ops as ops
        from stylegan2ada.torch_utils.ops import bias_act, conv2d_gradfix
        assert bias_act is ops.bias_act and conv2d_gradfix is ops.conv2d_gradfix
        import train_parts.generators as G
        from sgb200.modconv import modulated_conv2d
        assert G.modulated_conv2d is modulated_conv2d
        print('OK')
    ''', [os.path.join(ROOT, 'style-big-gan_b200', 'overlay'), REF])
    assert 'OK' in out
