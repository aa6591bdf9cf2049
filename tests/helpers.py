"""Shared helpers for the test-suite (golden file access, error metrics)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name))
        self.manifest = json.loads(str(self.z['manifest'])) if 'manifest' in self.z else None

    def cases(self):
        return self.manifest

    def t(self, cid, key, device='cpu', dtype=None):
        k = f'{cid}.{key}'
        if k not in self.z:
            return None
        v = torch.from_numpy(self.z[k]).to(device)
        return v if dtype is None else v.to(dtype)

    def raw(self, key):
        return self.z[key]


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the 'relative' of BASELINE.json's 1e-4 / 1e-2 tolerances."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    denom = max(b.abs().max().item(), 1e-12)
    return (a - b).abs().max().item() / denom


def assert_close(a, b, tol, what=''):
    e = rel_err(a, b)
    assert e <= tol, f'{what}: rel err {e:.3e} > {tol:.1e}'
