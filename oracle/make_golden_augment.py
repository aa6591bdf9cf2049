"""Golden vectors for the ADA AugmentPipe from the REAL reference on the CPU.

    python -m oracle.make_golden_augment        # rewrites tests/golden/augment_pipe.npz (authoring container only)

The reference's `train_parts/augmentations.py:121-432` AugmentPipe is the caller of three ops of the hot path in forms the
G / D networks never use: the 12-tap sym6 separable `upfirdn2d.upsample2d / downsample2d` (:294,305, 1-D filter, negative
padding, flip_filter), `grid_sample_gradfix.grid_sample` (:302) and the per-sample depthwise filter bank through
`conv2d_gradfix.conv2d(groups = N * C)` (:402-403).  With `debug_percentile` its random draws are replaced by fixed percentiles
of their distributions (:185 ff.) and with `p = 1` every probability mask is all-true, so the pipe is a deterministic function
of its input: the CPU result of the reference (impl='ref' ops) is a golden the GPU run of the UNCHANGED AugmentPipe on the
sgb200 ops (`sgb200.install()`) must reproduce (tests/test_ref_callers_gpu.py).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

from oracle.make_golden import _import_reference, OUT

# every deterministic branch on: blitting, geometric (12-tap up / grid_sample / 12-tap down), colour, image-space filter bank;
# noise and cutout draw from the device RNG and stay off
PIPE_KWARGS = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1,
                   saturation=1, imgfilter=1)
CASES = [dict(name='p70_64', res=64, batch=4, percentile=0.7, seed=11),
         dict(name='p25_48', res=48, batch=3, percentile=0.25, seed=12)]      # 48: not a power of two, odd padding algebra


def run_case(aug_mod, case, device='cpu'):
    """(images, out, d sum(out * w) / d images, w) of one case on `device` with the ops `aug_mod` is bound to"""
    g = torch.Generator().manual_seed(case['seed'])
    images = (torch.rand(case['batch'], 3, case['res'], case['res'], generator=g) * 2 - 1).to(device).requires_grad_(True)
    w = torch.randn(case['batch'], 3, case['res'], case['res'], generator=g).to(device)
    pipe = aug_mod.AugmentPipe(**PIPE_KWARGS).to(device)
    pipe.p.copy_(torch.ones([]))
    out = pipe(images, debug_percentile=case['percentile'])
    gi, = torch.autograd.grad((out * w).sum(), [images])
    return images.detach(), out.detach(), gi.detach(), w


def main():
    _import_reference()
    from train_parts import augmentations as aug_mod
    torch.set_num_threads(os.cpu_count() or 1)
    arrays = {}
    for case in CASES:
        images, out, gi, w = run_case(aug_mod, case)
        assert out.shape == images.shape and torch.isfinite(out).all() and torch.isfinite(gi).all()
        arrays[case['name'] + '.images'] = images.numpy()
        arrays[case['name'] + '.w'] = w.numpy()
        arrays[case['name'] + '.out'] = out.numpy()
        arrays[case['name'] + '.grad_images'] = gi.numpy()
        print(case['name'], 'out', tuple(out.shape), 'abs max', float(out.abs().max()), 'grad abs max', float(gi.abs().max()))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'augment_pipe.npz'), **arrays)
    print('wrote', os.path.join(OUT, 'augment_pipe.npz'))


if __name__ == '__main__':
    sys.exit(main())
