"""Snapshot the UNMODIFIED reference checkout into the git-ignored `baseline/_ref/` so that it travels to the GPU box
(`gpurun` ships everything that is not listed in `.gpurunignore`; `/root/reference` itself does not exist there).

    python baseline/snapshot_reference.py [--src /root/reference]

What it is for (nothing under `style-big-gan_b200/` imports it):
  * `-m gpu` tests that run the reference's OWN callers (`train_parts/generators.py`, `discriminators.py`,
    `losses_base.py`, `regularizations.py`) on the sgb200 kernels through `sgb200.install()` (tests/test_ref_callers_gpu.py);
  * the "kernel to beat" column: the reference's own CUDA path (its JIT `bias_act_plugin` / `upfirdn2d_plugin` +
    cuDNN through `F.conv2d`) timed on the same B200 (`benchmarks/op_sweep.py --vs-reference`, `bench.py`);
  * the reference arm `bench.py --impl reference` (`cpu_baseline.kind = "reference"`): the reference's own modules
    on the box's host cores (impl='ref' path).

The reference is not a Python package (no setup.py / pyproject.toml), so the `pip install --target baseline/_ref`
recipe of the task contract does not apply; this copy is the equivalent.  Only source files are copied (.py, .cu,
.cpp, .h, .yaml, .txt, .md), byte for byte; a MANIFEST.json with their sha256 is written next to them.
"""
import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
EXT = ('.py', '.cu', '.cpp', '.h', '.yaml', '.txt', '.md')


def snapshot(src='/root/reference', dst=DST, quiet=False):
    """Returns the number of files copied, or None when `src` does not exist (GPU box: the prebuilt copy is used)."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ('.git', '__pycache__')]
        for f in files:
            if not f.endswith(EXT):
                continue
            p = os.path.join(base, f)
            rel = os.path.relpath(p, src)
            q = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(q), exist_ok=True)
            with open(p, 'rb') as fh:
                data = fh.read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            if not (os.path.exists(q) and open(q, 'rb').read() == data):
                shutil.copyfile(p, q)
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as fh:
        json.dump(dict(source=src, files=manifest), fh, indent=1, sort_keys=True)
    if not quiet:
        print(f'baseline/_ref: {len(manifest)} reference source files from {src}')
    return len(manifest)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    a = ap.parse_args()
    if snapshot(a.src) is None:
        raise SystemExit(f'{a.src} does not exist')
