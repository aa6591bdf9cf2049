#!/usr/bin/env python
"""GPU check of the TMA-staged convolution kernel (csrc/conv_tma.cuh) against torch on the CPU-independent reference
(torch.nn.functional on the GPU in fp32 with TF32 off) and timing against the halo kernel.
    SGB_TMA_BO=0|1 python benchmarks/experiments/tma_check.py
Prints one line per case: max-norm relative error and median ms with the TMA kernel (SGB_TMA=1 in this process)."""
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from sgb200.ops import conv2d_gradfix as cg  # noqa: E402
from sgb200 import _lib  # noqa: E402

DEV = 'cuda'


def timeit(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    cases = [
        # (N, Ci, Co, H, W, k, pad, transposed, dtype, scale)
        (2, 32, 32, 64, 64, 3, 1, False, torch.float16, False),
        (2, 32, 32, 64, 64, 3, 1, False, torch.float16, True),
        (2, 64, 64, 48, 80, 3, 1, False, torch.float16, False),
        (2, 64, 64, 33, 47, 3, 1, False, torch.float16, True),
        (2, 128, 128, 40, 40, 3, 1, False, torch.float16, False),
        (2, 256, 256, 32, 32, 3, 1, False, torch.float16, True),
        (2, 128, 64, 64, 64, 3, 1, True, torch.float16, False),
        (2, 64, 64, 64, 64, 1, 0, False, torch.float16, False),
        (2, 40, 72, 37, 53, 3, 1, False, torch.float16, False),
        (2, 64, 64, 64, 64, 3, 1, False, torch.float32, False),
        (2, 64, 64, 64, 64, 3, 1, False, torch.float32, True),
        (2, 128, 128, 35, 35, 3, 1, True, torch.float32, False),
        (2, 16, 32, 64, 64, 3, 1, False, torch.float32, False),
        (2, 512, 512, 32, 32, 3, 1, False, torch.float32, False),
        (4, 32, 32, 1024, 1024, 3, 1, False, torch.float16, False),
        (4, 64, 64, 512, 512, 3, 1, False, torch.float16, False),
        (4, 128, 128, 256, 256, 3, 1, False, torch.float16, False),
        (32, 64, 64, 256, 256, 3, 1, False, torch.float32, False),
        (32, 128, 128, 128, 128, 3, 1, False, torch.float32, False),
        (32, 256, 256, 64, 64, 3, 1, False, torch.float32, True),
        (8, 512, 512, 64, 64, 3, 1, False, torch.float32, True),
        (8, 512, 512, 64, 64, 3, 1, True, torch.float32, False),
    ]
    for (n, ci, co, h, w, k, pad, tr, dtype, scale) in cases:
        x = torch.randn(n, ci, h, w, device=DEV, dtype=dtype).contiguous(memory_format=torch.channels_last)
        wshape = (ci, co, k, k) if tr else (co, ci, k, k)
        wt = (torch.randn(wshape, device=DEV) / math.sqrt(ci * k * k)).to(dtype)
        s = (torch.randn(n, ci, device=DEV) + 1) if scale else None
        op = cg.conv_transpose2d if tr else cg.conv2d
        fo = F.conv_transpose2d if tr else F.conv2d
        _lib.profile_start()
        with torch.no_grad():
            y = op(x, wt, padding=pad, in_scale=s)
        torch.cuda.synchronize()
        kinds = sorted(k.split(' k')[0][:12] + (' TMA' if k.endswith(' tma') else ' halo') for k in _lib.profile_stop().summary(by_tag=True))
        small = n * h * w <= 4 * 256 * 256
        err = float('nan')
        if small or True:
            torch.backends.cudnn.allow_tf32 = False
            with torch.no_grad():
                xs = x.float() * (s[:, :, None, None] if scale else 1.0)
                yo = fo(xs[:2], wt.float(), padding=pad)
            torch.backends.cudnn.allow_tf32 = True
            err = float((y[:2].float() - yo).abs().max() / yo.abs().max())
        with torch.no_grad():
            ms = timeit(lambda: op(x, wt, padding=pad, in_scale=s))
        flops = 2.0 * n * y.shape[2] * y.shape[3] * ci * co * k * k
        print(f'{str(dtype)[6:]:8s} x[{n},{ci},{h},{w}] co{co} k{k} {"T" if tr else " "} {"mod" if scale else "   "}  err {err:.2e}  {ms:.4f} ms  '
              f'{flops / ms / 1e9:7.1f} TF/s  {kinds}', flush=True)


if __name__ == '__main__':
    main()
