"""CPU: the C-ABI library loads and exports every symbol include/sgb200.h declares (no compute)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'sgb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sgb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_bound_and_exported():
    from sgb200 import _lib
    declared = _declared_symbols()
    assert declared, 'no declarations parsed'
    assert sorted(_lib.SIGNATURES) == declared
    lib = _lib.lib()                      # raises if the .so is missing or lacks a symbol
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sgb_abi_version() == 1
    assert lib.sgb_last_error() is not None


def test_no_cpu_fallback():
    import torch
    import sgb200
    with pytest.raises(RuntimeError):
        sgb200.ops.bias_act.bias_act(torch.zeros(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError):
        sgb200.ops.upfirdn2d.upfirdn2d(torch.zeros(1, 1, 4, 4), None)
    with pytest.raises(RuntimeError):
        sgb200.ops.conv2d_gradfix.conv2d(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 3, 3))
    with pytest.raises(RuntimeError):
        sgb200.modulated_conv2d(torch.zeros(1, 2, 4, 4), torch.zeros(2, 2, 3, 3), torch.ones(1, 2), fused_modconv=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'style-big-gan_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{fn} imports the oracle'
