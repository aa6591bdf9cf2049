"""CPU: the committed AugmentPipe golden (tests/golden/augment_pipe.npz) reproduces from the reference snapshot baseline/_ref
(the reference's own train_parts/augmentations.py on its impl='ref' ops).  Skipped where the snapshot is absent."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from benchmarks import ref_harness
ref_harness.import_reference('reference')
from oracle import make_golden_augment as MG
import train_parts.augmentations as aug_mod
z = np.load(os.path.join(%(root)r, 'tests', 'golden', 'augment_pipe.npz'))
for case in MG.CASES[1:]:          # the small case keeps the CPU suite fast
    images, out, gi, _ = MG.run_case(aug_mod, case)
    assert np.array_equal(images.numpy(), z[case['name'] + '.images'])
    assert np.abs(out.numpy() - z[case['name'] + '.out']).max() <= 1e-5 * np.abs(z[case['name'] + '.out']).max()
    assert np.abs(gi.numpy() - z[case['name'] + '.grad_images']).max() <= 1e-5 * np.abs(z[case['name'] + '.grad_images']).max()
print('AUGMENT_GOLDEN_OK')
'''


def test_augment_pipe_golden_reproduces_from_the_snapshot():
    if not os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'train_parts', 'augmentations.py')):
        pytest.skip('baseline/_ref snapshot missing')
    # own process: the reference backend binds the same module names sgb200.install() rebinds in other tests
    r = subprocess.run([sys.executable, '-c', CODE % dict(root=ROOT)], capture_output=True, text=True, timeout=600)
    assert 'AUGMENT_GOLDEN_OK' in r.stdout, r.stderr[-2000:]
