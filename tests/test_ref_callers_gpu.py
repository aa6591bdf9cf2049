"""GPU: the reference's UNCHANGED callers on the sgb200 kernels.

`benchmarks/ref_harness.py` puts the snapshot `baseline/_ref/` (a byte-for-byte copy of the reference checkout made by
`baseline/snapshot_reference.py`; it travels to the GPU box, `/root/reference` does not) on sys.path behind
`sgb200.install()`.  Everything that runs here above the ops is the reference's own code:
`train_parts/generators.py` ('sg2_classic' G: SynthesisLayer :310-329, ToRGBLayer :344-348, SynthesisBlock :414-458),
`train_parts/discriminators.py` (Conv2dLayer :115-124, DiscriminatorBlock :270-302), `train_parts/losses_base.py`
(SG2Loss :83-109,131-156) and `train_parts/regularizations.py` (PPLreg :11-37, R1reg :40-56).

Golden = `tests/golden/net_tiny.npz`, produced by the same reference classes on CPU (impl='ref'), all parameter gradients
of the four training phases (oracle/make_golden.py).  Modes:
  strict   torch.backends.cudnn.allow_tf32 = False  -> fp32 FFMA kernels                       every tensor <= 2e-4 / 5e-4 (2nd order)
  tf32     allow_tf32 = True -> tcgen05 kind::tf32 forward / dgrad / wgrad (bench.py default)
  fp16     num_fp16_res = 2, conv_clamp = 256 (+ TF32 for the fp32 blocks): config C / D numerics, against the fp32 golden
           (the clamp never engages at these magnitudes)
           tensor-core modes: median over tensors <= 1e-2, worst tensor <= 2 x the error the reference's own cuDNN path shows
           on the same gradients (helpers.REFERENCE_GPU_WORST, profiles/r2_parity_report.md)
Metric: per-tensor max-norm relative error (helpers.check_phase_grads).
"""
import contextlib
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close, check_phase_grads, phase_tolerances

pytestmark = pytest.mark.gpu
DEV = 'cuda'

MODES = {
    'strict': dict(tf32=False, num_fp16_res=0, conv_clamp=None, tolf=1e-4),
    'tf32': dict(tf32=True, num_fp16_res=0, conv_clamp=None, tolf=1e-2),
    'fp16': dict(tf32=True, num_fp16_res=2, conv_clamp=256, tolf=1e-2),
}


@pytest.fixture(scope='module')
def H():
    from benchmarks import ref_harness
    if ref_harness.reference_root() is None:
        pytest.skip('baseline/_ref snapshot missing (python baseline/snapshot_reference.py where /root/reference exists)')
    ref_harness.import_reference('sgb200')
    return ref_harness


@pytest.fixture(scope='module')
def gold():
    z = np.load(os.path.join(GOLDEN, 'net_tiny.npz'))
    return z, json.loads(str(z['meta']))


@contextlib.contextmanager
def _tf32(on):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = on
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@contextlib.contextmanager
def _fixed_randn_like(value):
    """PPLreg draws its noise with torch.randn_like(gen_img) (regularizations.py:24); the golden file holds the draw."""
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: value.to(device=t.device, dtype=t.dtype)
    try:
        yield
    finally:
        torch.randn_like = orig


def _tiny_trainer(H, meta, mode):
    c, m = meta['cfg'], MODES[mode]
    w = dict(res=c['img_resolution'], batch_gpu=meta['n'], z_dim=c['z_dim'], w_dim=c['w_dim'], map_layers=c['map_layers'],
             channel_base=c['channel_base'], d_arch=c['d_arch'], mbstd=c['mbstd_group_size'], r1_gamma=meta['r1_gamma'], ppl=True,
             style_mixing_prob=0.0, num_fp16_res=m['num_fp16_res'], conv_clamp=m['conv_clamp'], ema_kimg=10.0, g_attn=(), d_attn=())
    return H.RefCallerTrainer(w, DEV, 'sgb200', noise_mode='const', channel_max=c['channel_max'], use_ema=False)


def _load(z, G, D):
    G.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('G.')}, strict=True)
    D.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('D.')}, strict=True)


def test_reference_modules_are_bound_to_sgb200(H):
    import train_parts.generators as G
    import train_parts.discriminators as Dm
    import train_parts.regularizations as Rg
    import sgb200.ops as ops
    from sgb200.modconv import modulated_conv2d
    assert G.__file__.startswith(H.reference_root())
    assert G.bias_act is ops.bias_act and G.upfirdn2d is ops.upfirdn2d and G.conv2d_resample is ops.conv2d_resample
    assert Dm.bias_act is ops.bias_act and Dm.conv2d_resample is ops.conv2d_resample and Rg.conv2d_gradfix is ops.conv2d_gradfix
    assert G.modulated_conv2d is modulated_conv2d


@pytest.mark.parametrize('mode', list(MODES))
def test_reference_G_D_forward_on_sgb200(H, gold, mode):
    from sgb200 import _lib
    z, meta = gold
    m = MODES[mode]
    tr = _tiny_trainer(H, meta, mode)
    _load(z, tr.G, tr.D)
    zz = torch.from_numpy(z['z']).to(DEV)
    c = torch.zeros(meta['n'], 0, device=DEV)
    n0 = _lib.launch_count()
    with torch.no_grad(), _tf32(m['tf32']):
        ws = tr.G.mapping(zz, c, skip_w_avg_update=True)
        img = tr.G.synthesis(ws, noise_mode='const')
        logits = tr.D(img, c)
    assert _lib.launch_count() - n0 > 50, 'the reference callers did not reach libsgb200'
    assert_close(img, torch.from_numpy(z['img']), m['tolf'], f'img [{mode}]')
    assert_close(logits, torch.from_numpy(z['logits']), 2 * m['tolf'], f'logits [{mode}]')


@pytest.mark.parametrize('mode', list(MODES))
def test_reference_SG2Loss_R1_PPL_on_sgb200(H, gold, mode):
    """All parameter gradients of Gmain / Dmain / Dreg (R1) / Greg (PPL) computed by the reference's own
    SG2Loss.accumulate_gradients + R1reg + PPLreg on the sgb200 kernels, against the reference's CPU results."""
    z, meta = gold
    m = MODES[mode]
    tr = _tiny_trainer(H, meta, mode)
    _load(z, tr.G, tr.D)
    zz = torch.from_numpy(z['z']).to(DEV)
    real = torch.from_numpy(z['real']).to(DEV)
    pl_noise = torch.from_numpy(z['pl_noise'])
    gains = meta['gains']
    worst = {}
    with _tf32(m['tf32']):
        for phase, tag in [('Gmain', 'G.'), ('Dmain', 'D.'), ('Dreg', 'D.'), ('Greg', 'G.')]:
            with _fixed_randn_like(pl_noise):
                ph = tr.phase_grads(phase, real, zz, gains[phase])
            worst[phase] = check_phase_grads(z, phase, tag, ph.module, *phase_tolerances(mode, phase))
    print(f'[{mode}] (worst, median) per-tensor rel err: ' + ', '.join(f'{k} {v[0]:.2e}/{v[1]:.2e}' for k, v in worst.items()))


def test_reference_callers_tf32_reach_tensor_core_kernels(H, gold):
    """fp32 blocks of the reference are NCHW (generators.py:392): with allow_tf32 the convolutions must still run the
    tcgen05 kernels (conv2d_gradfix.channels_last_for_tensor_cores), not the SIMT fallback."""
    from sgb200 import _lib
    z, meta = gold
    tr = _tiny_trainer(H, meta, 'tf32')
    zz = torch.from_numpy(z['z']).to(DEV)
    real = torch.from_numpy(z['real']).to(DEV)
    with _tf32(True):
        tr.phase_grads('Gmain', real, zz, 1)          # warm
        _lib.profile_start()
        tr.phase_grads('Gmain', real, zz, 1)
        tr.phase_grads('Dreg', real, zz, 4)
        torch.cuda.synchronize()
        summ = _lib.profile_stop().summary()
    assert summ.get('conv_fwd_tc', {}).get('launches', 0) > 20 and summ.get('conv_wgrad_tc', {}).get('launches', 0) > 10, summ.keys()
    assert 'conv_wgrad_simt' not in summ, 'a weight gradient fell back to the SIMT kernel'


def test_reference_training_iterations_run(H):
    """A few whole iterations of the reference loop body (lazy reg schedule, nan_to_num, Adam, G_ema) at the
    sg2attent topology scaled down: attention blocks (biggan/layers.py:144-169) in G and D, fp16 + clamp."""
    w = dict(H.WORKLOADS['sg2attent256'], res=32, batch_gpu=4, z_dim=64, w_dim=64, channel_base=1024, mbstd=4, num_fp16_res=2,
             g_attn=(16, 8), d_attn=(16,))
    with _tf32(True):
        tr = H.RefCallerTrainer(w, DEV, 'sgb200', channel_max=64)
        real = torch.randint(0, 256, [4, 3, 32, 32], dtype=torch.uint8, device=DEV)
        p0 = [p.detach().clone() for p in tr.G.parameters()]
        seen = [tr.iteration(real) for _ in range(5)]
    torch.cuda.synchronize()
    assert seen[0] == ['Gmain', 'Dmain', 'Dreg'] and seen[1] == ['Gmain', 'Dmain'] and seen[4] == ['Gmain', 'Dmain', 'Dreg']
    assert all(torch.isfinite(p).all() for p in list(tr.G.parameters()) + list(tr.D.parameters()))
    assert any((a != b.detach()).any() for a, b in zip(p0, tr.G.parameters()))


# ---------------------------------------------------------------------------------------------------------------------
# ADA AugmentPipe: the reference's unchanged train_parts/augmentations.py:121-432 on the sgb200 ops.
# It reaches the ops in forms the networks never use: 12-tap sym6 separable upsample2d / downsample2d with negative padding and
# flip_filter (:294,305), grid_sample_gradfix (:302) and the per-sample depthwise filter bank conv2d(groups = N * C) (:402-403).
# With debug_percentile and p = 1 the pipe is deterministic; golden = the same class on the reference's CPU path
# (oracle/make_golden_augment.py -> tests/golden/augment_pipe.npz).
def test_reference_augment_pipe_on_sgb200(H):
    from oracle import make_golden_augment as MG
    import train_parts.augmentations as aug_mod
    import sgb200.ops as ops
    assert aug_mod.upfirdn2d is ops.upfirdn2d and aug_mod.conv2d_gradfix is ops.conv2d_gradfix
    z = np.load(os.path.join(GOLDEN, 'augment_pipe.npz'))
    from sgb200 import _lib
    with _tf32(False):
        for case in MG.CASES:
            n0 = _lib.launch_count()
            images, out, gi, _ = MG.run_case(aug_mod, case, device=DEV)
            assert _lib.launch_count() > n0, 'AugmentPipe did not reach libsgb200'
            assert np.array_equal(images.cpu().numpy(), z[case['name'] + '.images'])
            # bilinear grid_sample + 12-tap filters in fp32: the CPU and CUDA arithmetic differ in summation order only
            assert_close(out, torch.from_numpy(z[case['name'] + '.out']), 2e-4, f"AugmentPipe {case['name']} forward")
            assert_close(gi, torch.from_numpy(z[case['name'] + '.grad_images']), 5e-4, f"AugmentPipe {case['name']} d/d images")
