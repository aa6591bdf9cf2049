#!/usr/bin/env python
"""The "kernel to beat" column: every op of the hot path timed on the same B200,
same shapes, same inputs,
    ours       sgb200.ops.* (libsgb200)
    reference  the reference's own CUDA path from the snapshot baseline/_ref: `bias_act_plugin` / `upfirdn2d_plugin`
               JIT-built by its custom_ops.get_plugin (bias_act.cu / upfirdn2d.cu compiled for sm_100a), cuDNN through
               torch.nn.functional.conv2d / conv_transpose2d (what conv2d_gradfix.py:38,43 resolve to on torch >= 2.0),
               and its modulated_conv2d (train_parts/generators.py:42-100) on top of those.
Both in ONE process (the sgb200 ops are imported under their own names; nothing is installed over the reference).

Timing: CUDA events, 5 warm-up + median of 20, a 256 MB buffer is overwritten between timed launches (L2 flush).
cuDNN runs with torch.backends.cudnn.benchmark = True (its best algorithm per shape) and the same allow_tf32 as ours.
Layouts: ours channels_last (its fast path); the reference is timed in BOTH layouts and the better one is reported
(the reference keeps fp32 blocks NCHW and fp16 blocks channels_last).

    python benchmarks/vs_reference.py [--quick] [--out profiles/r2_vs_reference.jsonl] [--md profiles/r2_vs_reference.md]
"""
import argparse
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default=None)
    ap.add_argument('--md', default=None)
    args = ap.parse_args()
    from benchmarks import ref_harness
    ref_harness.import_reference('reference')
    from stylegan2ada.torch_utils.ops import bias_act as r_bias_act, upfirdn2d as r_upfirdn2d
    import train_parts.generators as r_gen
    from sgb200.ops import bias_act, upfirdn2d, conv2d_gradfix
    from sgb200 import modulated_conv2d
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    rows = []

    def emit(op, shape, dtype, ours, ref, ref_layout):
        r = dict(op=op, shape=shape, dtype=dtype, ours_ms=round(ours, 4), reference_ms=round(ref, 4), reference_layout=ref_layout,
                 speedup=round(ref / ours, 3))
        rows.append(r)
        print(json.dumps(r), flush=True)

    def cl(t):
        return t.contiguous(memory_format=torch.channels_last)

    def best_ref(fn_of_layout):
        """time the reference in NCHW and channels_last, return (ms, layout) of the better"""
        res = {}
        for name, conv in (('nchw', lambda t: t.contiguous()), ('channels_last', cl)):
            try:
                res[name] = timeit(fn_of_layout(conv))
            except Exception as e:       # a layout the reference op refuses
                res[name] = float('inf')
                print(f'# reference failed in {name}: {type(e).__name__}: {e}', flush=True)
        k = min(res, key=res.get)
        return res[k], k

    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(DEV)
    # (resolution, channels, batch): layers of config B (ffhq256, N = 32) and config C (f1024, N = 4)
    layers = [(4, 512, 32), (16, 512, 32), (32, 512, 32), (64, 256, 32), (128, 128, 32), (256, 64, 32),
              (64, 512, 4), (128, 256, 4), (256, 128, 4), (512, 64, 4), (1024, 32, 4)]
    if args.quick:
        layers = [(32, 512, 32), (256, 64, 32), (512, 64, 4)]
    for res, c, n in layers:
        for dtype in ((torch.float32,) if n == 32 else (torch.float16,) if res >= 128 else (torch.float32,)):
            name = str(dtype).replace('torch.', '') + ('(tf32)' if dtype == torch.float32 else '')
            shape = [n, c, res, res]
            x0 = torch.randn(n, c, res, res, device=DEV, dtype=dtype)
            x = cl(x0)
            w32 = torch.randn(c, c, 3, 3, device=DEV) / math.sqrt(9 * c)
            w = w32.to(dtype)
            s = torch.randn(n, c, device=DEV) + 1
            b = torch.randn(c, device=DEV, dtype=dtype)
            with torch.no_grad():
                # --- plain convolutions: forward, data gradient (= transposed conv), stride-2 forms
                ours = timeit(lambda: conv2d_gradfix.conv2d(x, w, padding=1))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0), ww=conv(w): F.conv2d(xx, ww, padding=1)))
                emit('conv3x3 fwd', shape, name, ours, ref, lay)
                ours = timeit(lambda: conv2d_gradfix.conv_transpose2d(x, w, padding=1))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0), ww=conv(w): F.conv_transpose2d(xx, ww, padding=1)))
                emit('conv3x3 dgrad', shape, name, ours, ref, lay)
                if res <= 512:
                    wt = w.transpose(0, 1).contiguous()
                    ours = timeit(lambda: conv2d_gradfix.conv_transpose2d(x, wt, stride=2))
                    ref, lay = best_ref(lambda conv: (lambda xx=conv(x0), ww=conv(wt): F.conv_transpose2d(xx, ww, stride=2)))
                    emit('convT3x3 stride2 (up path)', shape, name, ours, ref, lay)
                if res >= 8:
                    xs0 = torch.randn(n, c, res + 1, res + 1, device=DEV, dtype=dtype)
                    xs = cl(xs0)
                    ours = timeit(lambda: conv2d_gradfix.conv2d(xs, w, stride=2))
                    ref, lay = best_ref(lambda conv: (lambda xx=conv(xs0), ww=conv(w): F.conv2d(xx, ww, stride=2)))
                    emit('conv3x3 stride2 (down path)', [n, c, res + 1, res + 1], name, ours, ref, lay)
                    del xs0, xs
                # --- memory-bound ops against the reference's plugins
                ours = timeit(lambda: bias_act.bias_act(x, b, act='lrelu', gain=math.sqrt(2), clamp=256))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_bias_act.bias_act(xx, b, act='lrelu', gain=math.sqrt(2), clamp=256)))
                emit('bias_act fwd', shape, name, ours, ref, lay)
                ours = timeit(lambda: upfirdn2d.upfirdn2d(x, f, padding=[2, 1, 2, 1]))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_upfirdn2d.upfirdn2d(xx, f, padding=[2, 1, 2, 1])))
                emit('upfirdn2d FIR pad', shape, name, ours, ref, lay)
                ours = timeit(lambda: upfirdn2d.downsample2d(x, f))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_upfirdn2d.downsample2d(xx, f)))
                emit('upfirdn2d down2', shape, name, ours, ref, lay)
                if res <= 512:
                    ours = timeit(lambda: upfirdn2d.upsample2d(x, f))
                    ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_upfirdn2d.upsample2d(xx, f)))
                    emit('upfirdn2d up2', shape, name, ours, ref, lay)
                # --- modulated convolution (training form: fused_modconv=False), forward
                ours = timeit(lambda: modulated_conv2d(x, w32, s, padding=1, fused_modconv=False))
                ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_gen.modulated_conv2d(xx, w32, s, padding=1, fused_modconv=False)))
                emit('modulated_conv2d 3x3 fwd', shape, name, ours, ref, lay)
                if res <= 512:
                    ours = timeit(lambda: modulated_conv2d(x, w32, s, up=2, padding=1, resample_filter=f, flip_weight=False, fused_modconv=False))
                    ref, lay = best_ref(lambda conv: (lambda xx=conv(x0): r_gen.modulated_conv2d(xx, w32, s, up=2, padding=1, resample_filter=f,
                                                                                               flip_weight=False, fused_modconv=False)))
                    emit('modulated_conv2d 3x3 up2 fwd', shape, name, ours, ref, lay)
            # --- weight gradient, bias_act backward, modconv forward+backward (through autograd)
            wg = w.detach().clone().requires_grad_(True)
            y = conv2d_gradfix.conv2d(x, wg, padding=1)
            dy = torch.randn_like(y)
            ours = timeit(lambda: torch.autograd.grad(y, [wg], dy, retain_graph=True))

            def ref_wgrad(conv):
                xx, wr = conv(x0), conv(w).detach().requires_grad_(True)
                yr = F.conv2d(xx, wr, padding=1)
                dyr = conv(dy)
                return lambda: torch.autograd.grad(yr, [wr], dyr, retain_graph=True)
            ref, lay = best_ref(ref_wgrad)
            emit('conv3x3 wgrad', shape, name, ours, ref, lay)
            xb = x.detach().requires_grad_(True)
            yb = bias_act.bias_act(xb, b, act='lrelu', gain=math.sqrt(2), clamp=256)
            ours = timeit(lambda: torch.autograd.grad(yb, [xb], dy, retain_graph=True))

            def ref_bbwd(conv):
                xr = conv(x0).detach().requires_grad_(True)
                yr = r_bias_act.bias_act(xr, b, act='lrelu', gain=math.sqrt(2), clamp=256)
                dyr = conv(dy)
                return lambda: torch.autograd.grad(yr, [xr], dyr, retain_graph=True)
            ref, lay = best_ref(ref_bbwd)
            emit('bias_act bwd (dx)', shape, name, ours, ref, lay)

            def fwd_bwd(mc, xin):
                xr = xin.detach().requires_grad_(True)
                wr = w32.detach().requires_grad_(True)
                sr = s.detach().requires_grad_(True)

                def run():
                    out = mc(xr, wr, sr, padding=1, fused_modconv=False)
                    torch.autograd.grad(out, [xr, wr, sr], dy.to(out.dtype) if out.shape == dy.shape else torch.ones_like(out))
                return run
            ours = timeit(fwd_bwd(modulated_conv2d, x))
            ref, lay = best_ref(lambda conv: fwd_bwd(r_gen.modulated_conv2d, conv(x0)))
            emit('modulated_conv2d 3x3 fwd+bwd', shape, name, ours, ref, lay)
            del x, x0, y, dy, xb, yb
            torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, 'w') as fh:
            for r in rows:
                fh.write(json.dumps(r) + '\n')
    if args.md:
        with open(args.md, 'w') as fh:
            fh.write('| op | x shape | dtype | sgb200 ms | reference GPU ms (best layout) | reference / sgb200 |\n|---|---|---|---|---|---|\n')
            for r in rows:
                fh.write(f"| {r['op']} | {r['shape']} | {r['dtype']} | {r['ours_ms']} | {r['reference_ms']} ({r['reference_layout']}) | "
                         f"{r['speedup']} |\n")


if __name__ == '__main__':
    main()
