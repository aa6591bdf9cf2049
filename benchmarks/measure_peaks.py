#!/usr/bin/env python
"""Measure the tensor-core rates the roofline fractions are quoted against, on this GPU, the way MEASURED_PEAKS.json was made
(torch.matmul 8192^3: best of 10 = burst, back to back for ~3 s = sustained), for the arithmetic types the convolutions use:
TF32 (fp32 tensors with torch.backends.cuda.matmul.allow_tf32), fp16 and bf16.  Writes one JSON object (stdout, or argv[1]).

    python benchmarks/measure_peaks.py [out.json]
"""
import json
import sys
import time

import torch


def rate(dtype, tf32, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device='cuda', dtype=dtype)
    b = torch.randn(n, n, device='cuda', dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); reps = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12
    return dict(burst_tflops=best, sustained_tflops=sustained)


def main():
    out = dict(gpu=torch.cuda.get_device_name(0), how='torch.matmul 8192^3 (cuBLAS), best of 10 = burst, back to back for 3 s = sustained',
               tf32=rate(torch.float32, True), fp16=rate(torch.float16, False), bf16=rate(torch.bfloat16, False),
               fp32_simt=rate(torch.float32, False, n=4096))
    s = json.dumps(out, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(s + '\n')
    print(s)


if __name__ == '__main__':
    main()
