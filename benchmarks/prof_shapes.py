#!/usr/bin/env python
"""A handful of hot-path launches at fixed shapes, each run a few times: the target of `ncu --set full`
(see profiles/README.md).  Prints CUDA-event times per case when run without ncu.

    python benchmarks/prof_shapes.py [--cases a,b,...] [--reps 3]
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
import torch  # noqa: E402

from sgb200.ops import bias_act, upfirdn2d, conv2d_gradfix  # noqa: E402
from sgb200 import modulated_conv2d  # noqa: E402

DEV = 'cuda'


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default='all')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--inner', type=int, default=1, help='back-to-back calls per timed region (amortises launch overhead)')
    ap.add_argument('--graph', action='store_true', help='capture the inner calls in a CUDA graph and time its replay (device time only)')
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = True
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(DEV)

    def conv_case(n, c, res, dtype, co=None):
        co = co or c
        x = cl(torch.randn(n, c, res, res, device=DEV, dtype=dtype))
        w = (torch.randn(co, c, 3, 3, device=DEV) / math.sqrt(9 * c)).to(dtype)
        return x, w

    cases = {}

    def case(name):
        def deco(fn):
            cases[name] = fn
            return fn
        return deco

    @case('fwd_f32_c64_256')
    def _():
        x, w = conv_case(8, 64, 256, torch.float32)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 8 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('fused_act_f32_c64_256')
    def _():
        from sgb200.ops import fused_conv
        x, w = conv_case(8, 64, 256, torch.float32)
        b = torch.randn(64, device=DEV)
        return (lambda: fused_conv.conv2d_bias_act(x, w, b, padding=1, act='lrelu', clamp=256)), 2.0 * 8 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('fused_mod_f32_c64_256')
    def _():
        from sgb200.ops import fused_conv
        x, w = conv_case(8, 64, 256, torch.float32)
        b = torch.randn(64, device=DEV)
        s = torch.randn(8, 64, device=DEV) + 1
        dc = torch.rand(8, 64, device=DEV) + 0.5
        nz = torch.randn(8, 1, 256, 256, device=DEV)
        return (lambda: fused_conv.conv2d_bias_act(x, w, b, padding=1, styles=s, dcoefs=dc, noise=nz, act='lrelu', clamp=256)), \
            2.0 * 8 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('scaled_f32_c64_256')
    def _():
        x, w = conv_case(8, 64, 256, torch.float32)
        s = torch.randn(8, 64, device=DEV) + 1
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1, in_scale=s)), 2.0 * 8 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('fwd_f32_c128_128')
    def _():
        x, w = conv_case(8, 128, 128, torch.float32)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 8 * 128 * 128 * 128 * 128 * 9, 2 * x.numel() * 4

    @case('fwd_f16_c512_64')
    def _():
        x, w = conv_case(8, 512, 64, torch.float16)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 8 * 64 * 64 * 512 * 512 * 9, 2 * x.numel() * 2

    @case('fwd_f32_c512_32')
    def _():
        x, w = conv_case(32, 512, 32, torch.float32)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 32 * 32 * 32 * 512 * 512 * 9, 2 * x.numel() * 4

    @case('wgrad_f32_c64_256')
    def _():
        x, w = conv_case(8, 64, 256, torch.float32)
        w.requires_grad_(True)
        y = conv2d_gradfix.conv2d(x, w, padding=1)
        dy = torch.randn_like(y)
        return (lambda: torch.autograd.grad(y, [w], dy, retain_graph=True)), 2.0 * 8 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('wgrad_f32_c512_32')
    def _():
        x, w = conv_case(32, 512, 32, torch.float32)
        w.requires_grad_(True)
        y = conv2d_gradfix.conv2d(x, w, padding=1)
        dy = torch.randn_like(y)
        return (lambda: torch.autograd.grad(y, [w], dy, retain_graph=True)), 2.0 * 32 * 32 * 32 * 512 * 512 * 9, 2 * x.numel() * 4

    @case('convT_s2_f32_c128_128')
    def _():
        x = cl(torch.randn(8, 128, 128, 128, device=DEV))
        w = torch.randn(128, 64, 3, 3, device=DEV) / math.sqrt(9 * 128)
        return (lambda: conv2d_gradfix.conv_transpose2d(x, w, stride=2)), 2.0 * 8 * 128 * 128 * 128 * 64 * 9, (x.numel() + 8 * 64 * 257 * 257) * 4

    @case('conv_s2_f32_c64_256')
    def _():
        x = cl(torch.randn(8, 64, 257, 257, device=DEV))
        w = torch.randn(128, 64, 3, 3, device=DEV) / math.sqrt(9 * 64)
        return (lambda: conv2d_gradfix.conv2d(x, w, stride=2)), 2.0 * 8 * 128 * 128 * 128 * 64 * 9, (x.numel() + 8 * 128 * 128 * 128) * 4

    @case('fir_f16_c128_256')
    def _():
        x = cl(torch.randn(8, 128, 256, 256, device=DEV, dtype=torch.float16))
        return (lambda: upfirdn2d.upfirdn2d(x, f, padding=[2, 1, 2, 1])), 0, 2 * x.numel() * 2

    @case('up2_f16_c128_128')
    def _():
        x = cl(torch.randn(8, 128, 128, 128, device=DEV, dtype=torch.float16))
        return (lambda: upfirdn2d.upsample2d(x, f)), 0, 5 * x.numel() * 2

    @case('down2_f32_c64_256')
    def _():
        x = cl(torch.randn(8, 64, 256, 256, device=DEV))
        return (lambda: upfirdn2d.downsample2d(x, f)), 0, 1.25 * x.numel() * 4

    @case('bias_act_f16_c128_256')
    def _():
        x = cl(torch.randn(8, 128, 256, 256, device=DEV, dtype=torch.float16))
        b = torch.randn(128, device=DEV, dtype=torch.float16)
        return (lambda: bias_act.bias_act(x, b, act='lrelu', gain=math.sqrt(2), clamp=256)), 0, 2 * x.numel() * 2

    @case('modconv_up2_f32_c128_128')
    def _():
        x = cl(torch.randn(8, 128, 128, 128, device=DEV))
        w = torch.randn(64, 128, 3, 3, device=DEV)
        s = torch.randn(8, 128, device=DEV) + 1
        return (lambda: modulated_conv2d(x, w, s, up=2, padding=1, resample_filter=f, flip_weight=False, fused_modconv=False)), \
            2.0 * 8 * 128 * 128 * 128 * 64 * 9, (x.numel() + 8 * 64 * 256 * 256) * 4

    # config C (f1024) shapes: batch 4, fp16, few channels at 512^2 / 1024^2 (HBM-bound layers)
    @case('fwd_f16_c32_1024')
    def _():
        x, w = conv_case(4, 32, 1024, torch.float16)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 4 * 1024 * 1024 * 32 * 32 * 9, 2 * x.numel() * 2

    @case('fwd_f16_c64_512')
    def _():
        x, w = conv_case(4, 64, 512, torch.float16)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 4 * 512 * 512 * 64 * 64 * 9, 2 * x.numel() * 2

    @case('fwd_f32_c64_256_n32')
    def _():
        x, w = conv_case(32, 64, 256, torch.float32)
        return (lambda: conv2d_gradfix.conv2d(x, w, padding=1)), 2.0 * 32 * 256 * 256 * 64 * 64 * 9, 2 * x.numel() * 4

    @case('wgrad_f16_c32_1024')
    def _():
        x, w = conv_case(4, 32, 1024, torch.float16)
        w.requires_grad_(True)
        y = conv2d_gradfix.conv2d(x, w, padding=1)
        dy = torch.randn_like(y)
        return (lambda: torch.autograd.grad(y, [w], dy, retain_graph=True)), 2.0 * 4 * 1024 * 1024 * 32 * 32 * 9, 2 * x.numel() * 2

    @case('wgrad_f16_c64_512')
    def _():
        x, w = conv_case(4, 64, 512, torch.float16)
        w.requires_grad_(True)
        y = conv2d_gradfix.conv2d(x, w, padding=1)
        dy = torch.randn_like(y)
        return (lambda: torch.autograd.grad(y, [w], dy, retain_graph=True)), 2.0 * 4 * 512 * 512 * 64 * 64 * 9, 2 * x.numel() * 2

    @case('fir_f16_c32_1024')
    def _():
        x = cl(torch.randn(4, 32, 1025, 1025, device=DEV, dtype=torch.float16))
        return (lambda: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)), 0, 2 * x.numel() * 2

    @case('fir_f32_c64_256')
    def _():
        x = cl(torch.randn(32, 64, 257, 257, device=DEV))
        return (lambda: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)), 0, 2 * x.numel() * 4

    @case('down2_f16_c32_1024')
    def _():
        x = cl(torch.randn(4, 32, 1024, 1024, device=DEV, dtype=torch.float16))
        return (lambda: upfirdn2d.downsample2d(x, f)), 0, 1.25 * x.numel() * 2

    @case('up2_f32_c64_128')
    def _():
        x = cl(torch.randn(32, 64, 128, 128, device=DEV))
        return (lambda: upfirdn2d.upsample2d(x, f)), 0, 5 * x.numel() * 4

    # the 512-channel layers of config-f at 4 images per GPU (fp32, 4^2 ... 64^2): a handful of tiles, latency-bound
    def small_case(name, n, res, stride=1, transposed=False):
        @case(name)
        def _():
            if transposed:
                x = cl(torch.randn(n, 512, res, res, device=DEV))
                w = torch.randn(512, 512, 3, 3, device=DEV) / math.sqrt(9 * 512)
                oh = (res - 1) * stride + 3 - (2 if stride == 1 else 0)
                fn = (lambda: conv2d_gradfix.conv_transpose2d(x, w, stride=stride, padding=(1 if stride == 1 else 0)))
                return fn, 2.0 * n * res * res * 512 * 512 * 9, (x.numel() + n * 512 * oh * oh) * 4
            xin = res * stride + (1 if stride == 2 else 0)
            x = cl(torch.randn(n, 512, xin, xin, device=DEV))
            w = torch.randn(512, 512, 3, 3, device=DEV) / math.sqrt(9 * 512)
            fn = (lambda: conv2d_gradfix.conv2d(x, w, stride=stride, padding=(1 if stride == 1 else 0)))
            return fn, 2.0 * n * res * res * 512 * 512 * 9, (x.numel() + n * 512 * res * res) * 4
    for n_ in (4, 32):
        small_case(f'small_s1_8_n{n_}', n_, 8)
        small_case(f'small_s1_16_n{n_}', n_, 16)
        small_case(f'small_s1_32_n{n_}', n_, 32)
        small_case(f'small_s2_8_n{n_}', n_, 8, stride=2)          # x[n,512,17,17] -> 8 x 8
        small_case(f'small_s2_16_n{n_}', n_, 16, stride=2)
        small_case(f'small_T1_16_n{n_}', n_, 16, transposed=True)
        small_case(f'small_T2_8_n{n_}', n_, 8, stride=2, transposed=True)     # 8 x 8 -> 17 x 17

    names = list(cases) if args.cases == 'all' else args.cases.split(',')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name in names:
        with torch.no_grad() if not name.startswith('wgrad') else torch.enable_grad():
            fn, flops, nbytes = cases[name]()
        ts = []
        graph = None
        if args.graph and not name.startswith('wgrad'):
            with torch.no_grad():
                fn()                                   # warm-up: attributes, allocator
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for _ in range(args.inner):
                        fn()
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.inner if graph is None else 0):
                if name.startswith('wgrad'):
                    fn()
                else:
                    with torch.no_grad():
                        fn()
            if graph is not None:
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / args.inner)
        ms = min(ts)
        msg = f'{name:28s} ms={ms:.4f}'
        if flops:
            msg += f' tflops={flops / ms / 1e9:.1f}'
        msg += f' gbs={nbytes / ms / 1e6:.0f}'
        print(msg, flush=True)
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
