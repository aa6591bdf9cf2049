#!/usr/bin/env python
"""Op micro-benchmark sweep (BASELINE.json configs[4]): modulated_conv2d (3x3, 3x3 up=2, 1x1), plain conv
forward / dgrad / wgrad, upfirdn2d (up=2 / down=2 / FIR-only, [1,3,3,1]) and bias_act (lrelu, gain sqrt2, clamp 256)
across resolutions and channel widths, channels_last, fp16 and fp32 (TF32).

Timing: CUDA events on the current stream, 5 warm-up + median of 20, a 256 MB buffer is overwritten between
timed launches to flush the 126 MB L2.  Prints one JSON object per line:
  {"op", "shape", "dtype", "ms", "tflops" | "gbs", "frac_of_peak"}
Fractions are against MEASURED_PEAKS.json (HBM copy GB/s; bf16 dense TFLOP/s burst -- TF32 runs at half that rate).
"""
import argparse
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
import torch  # noqa: E402

from sgb200.ops import bias_act, upfirdn2d, conv2d_gradfix, conv2d_resample  # noqa: E402
from sgb200 import modulated_conv2d  # noqa: E402

DEV = 'cuda'
_flush = None


def timeit(fn, reps=20, warm=5):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d['bf16_tflops']
    return 6650.0, 1590.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = True
    hbm, tc = peaks()
    rows = []

    def emit(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(DEV)
    # (resolution, channels) pairs that occur in config B (cb 16384) and config C (cb 32768)
    layers = [(4, 512), (8, 512), (16, 512), (32, 512), (64, 512), (64, 256), (128, 256), (128, 128), (256, 128), (256, 64),
              (512, 64), (1024, 32)]
    if args.quick:
        layers = [(16, 512), (64, 256), (256, 64)]
    for dtype in (torch.float16, torch.float32):
        es = 2 if dtype == torch.float16 else 4
        tcp = tc if dtype == torch.float16 else tc / 2
        for res, c in layers:
            n = 8 if res < 512 else 4
            if n * c * res * res * es > (3 << 30):
                continue
            x = torch.randn(n, c, res, res, device=DEV, dtype=dtype).contiguous(memory_format=torch.channels_last)
            w = (torch.randn(c, c, 3, 3, device=DEV) / math.sqrt(9 * c))
            wd = w.to(dtype)
            s = torch.randn(n, c, device=DEV) + 1
            b = torch.randn(c, device=DEV, dtype=dtype)
            name = f'{dtype}'.replace('torch.', '')
            shape = [n, c, res, res]
            flops = 2.0 * n * res * res * c * c * 9
            conv_bytes = (2 * x.numel() + w.numel()) * es

            def rec(op, ms, fl=None, by=None):
                d = dict(op=op, shape=shape, dtype=name, ms=round(ms, 4))
                if fl:
                    d['tflops'] = round(fl / ms / 1e9, 2)
                    d['frac_tc'] = round(fl / ms / 1e9 / tcp, 4)
                if by:
                    d['gbs'] = round(by / ms / 1e6, 1)
                    d['frac_hbm'] = round(by / ms / 1e6 / hbm, 4)
                emit(**d)

            with torch.no_grad():
                rec('conv3x3_fwd', timeit(lambda: conv2d_gradfix.conv2d(x, wd, padding=1)), flops, conv_bytes)
                rec('conv3x3_dgrad', timeit(lambda: conv2d_gradfix.conv_transpose2d(x, wd, padding=1)), flops, conv_bytes)
                rec('modconv3x3', timeit(lambda: modulated_conv2d(x, w, s, padding=1, fused_modconv=False)), flops, conv_bytes)
                rec('bias_act_fwd', timeit(lambda: bias_act.bias_act(x, b, act='lrelu', gain=math.sqrt(2), clamp=256)), None, 2 * x.numel() * es)
                rec('fir_pad1', timeit(lambda: upfirdn2d.upfirdn2d(x, f, padding=[2, 1, 2, 1])), None, 2 * x.numel() * es)
                rec('downsample2', timeit(lambda: upfirdn2d.downsample2d(x, f)), None, 1.25 * x.numel() * es)
                if res <= 256:
                    rec('upsample2', timeit(lambda: upfirdn2d.upsample2d(x, f)), None, 5 * x.numel() * es)
                    rec('modconv3x3_up2', timeit(lambda: modulated_conv2d(x, w, s, up=2, padding=1, resample_filter=f, flip_weight=False,
                                                                          fused_modconv=False)), flops, (5 * x.numel() + w.numel()) * es)
            # wgrad + bias_act backward through autograd
            xg = x.detach().requires_grad_(False)
            wg = wd.detach().requires_grad_(True)
            y = conv2d_gradfix.conv2d(xg, wg, padding=1)
            dy = torch.randn_like(y)
            rec('conv3x3_wgrad', timeit(lambda: torch.autograd.grad(y, [wg], dy, retain_graph=True)), flops, conv_bytes)
            xb = x.detach().requires_grad_(True)
            yb = bias_act.bias_act(xb, b, act='lrelu', gain=math.sqrt(2), clamp=256)
            rec('bias_act_bwd', timeit(lambda: torch.autograd.grad(yb, [xb], dy, retain_graph=True)), None, 3 * x.numel() * es)
            del x, y, dy, xb, yb
            torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, 'w') as fh:
            for r in rows:
                fh.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
