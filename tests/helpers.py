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


def check_phase_grads(z, phase, tag, module, tol):
    """Parameter gradients of one training phase against the reference-generated golden file (net_tiny.npz).

    Metric = per-tensor max-norm relative error: max|got - ref| / max|ref| (NOT elementwise).  Tensors whose reference
    gradient is more than 100x smaller than the phase's largest one (e.g. D biases under R1: they only receive
    second-order signal through the minibatch-stddev layer, ~1e-8 against 1e-2 for the weights) are measured against
    that floor instead of their own tiny norm, where fp32 summation order alone exceeds any relative tolerance."""
    keys = [k for k in z.files if k.startswith(f'{phase}.grad.{tag}')]
    assert keys
    named = dict(module.named_parameters())
    floor = 1e-2 * max(float(np.abs(z[k]).max()) for k in keys)
    worst = 0.0
    for k in keys:
        name = k[len(f'{phase}.grad.{tag}'):]
        assert named[name].grad is not None, f'{phase}: no grad for {name}'
        ref = torch.from_numpy(z[k])
        got = named[name].grad.detach().float().cpu()
        assert got.shape == ref.shape
        err = (got.double() - ref.double()).abs().max().item() / max(ref.abs().max().item(), floor)
        assert err <= tol, f'{phase} {name}: rel err {err:.3e} > {tol:.1e}'
        worst = max(worst, err)
    return worst
