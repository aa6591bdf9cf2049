"""G / D networks on the sgb200 ops against the golden vectors of the reference networks (net_tiny.npz)
and against the CPU oracle.  CPU part: parameter naming; GPU part: forward + the four training phases."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close
from oracle import ref_networks as RN

DEV = 'cuda'


@pytest.fixture(scope='module')
def gold():
    z = np.load(os.path.join(GOLDEN, 'net_tiny.npz'))
    meta = json.loads(str(z['meta']))
    return z, meta


def _build(meta, device, **over):
    from sgb200 import training
    c = meta['cfg']
    cfg = training.TrainConfig(img_resolution=c['img_resolution'], z_dim=c['z_dim'], w_dim=c['w_dim'],
                               channel_base=c['channel_base'], channel_max=c['channel_max'], map_layers=c['map_layers'],
                               d_arch=c['d_arch'], mbstd_group_size=c['mbstd_group_size'], batch_gpu=meta['n'],
                               style_mixing_prob=0.0, noise_mode='const', use_ema=False, r1_gamma=meta['r1_gamma'],
                               pl_weight=meta['pl_weight'], pl_decay=meta['pl_decay'], **over)
    G, D = training.build_networks(cfg, device)
    return cfg, G, D


def _load(z, G, D):
    gsd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('G.')}
    dsd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('D.')}
    G.load_state_dict(gsd, strict=True)
    D.load_state_dict(dsd, strict=True)


def test_state_dict_names_match_reference(gold):
    z, meta = gold
    cfg, G, D = _build(meta, 'cpu')
    gref = {k[2:]: tuple(z[k].shape) for k in z.files if k.startswith('G.')}
    dref = {k[2:]: tuple(z[k].shape) for k in z.files if k.startswith('D.')}
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == gref
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == dref
    _load(z, G, D)


@pytest.mark.gpu
@pytest.mark.parametrize('channels_last', [False, True])
def test_forward_matches_reference_golden(gold, channels_last):
    z, meta = gold
    cfg, G, D = _build(meta, DEV, channels_last=channels_last)
    _load(z, G, D)
    G.train(); D.train()
    zz = torch.from_numpy(z['z']).to(DEV)
    with torch.no_grad():
        ws = G.mapping(zz, None, skip_w_avg_update=True)
        img = G.synthesis(ws, noise_mode='const')
        logits = D(img, None)
        G.eval()
        img_fused = G.synthesis(ws, noise_mode='const')       # eval => fused_modconv (grouped conv) branch
    assert_close(ws, torch.from_numpy(z['ws']), 1e-4, 'ws')
    assert_close(img, torch.from_numpy(z['img']), 1e-4, 'img')
    assert_close(img_fused, torch.from_numpy(z['img']), 1e-4, 'img (fused_modconv)')
    assert_close(logits, torch.from_numpy(z['logits']), 1e-4, 'logits')


def _check(z, phase, tag, module, tol):
    """Per-tensor max-norm relative error.  Tensors whose reference gradient is more than 100x smaller than
    the phase's largest one (e.g. D biases under R1: they only receive second-order signal through the
    minibatch-stddev layer, ~1e-8 against 1e-2 for the weights) are measured against that floor instead of
    their own tiny norm, where fp32 summation order alone exceeds any relative tolerance."""
    keys = [k for k in z.files if k.startswith(f'{phase}.grad.{tag}')]
    assert keys
    named = dict(module.named_parameters())
    floor = 1e-2 * max(float(np.abs(z[k]).max()) for k in keys)
    for k in keys:
        name = k[len(f'{phase}.grad.{tag}'):]
        assert named[name].grad is not None, f'{phase}: no grad for {name}'
        ref = torch.from_numpy(z[k])
        got = named[name].grad.detach().cpu()
        assert got.shape == ref.shape
        err = (got.double() - ref.double()).abs().max().item() / max(ref.abs().max().item(), floor)
        assert err <= tol, f'{phase} {name}: rel err {err:.3e} > {tol:.1e}'


@pytest.mark.gpu
@pytest.mark.parametrize('channels_last', [False, True])
def test_training_phases_match_reference_golden(gold, channels_last):
    """Gradients of every parameter for Gmain, Dmain, Dreg (R1) and Greg (path length) -- the last two are
    double backward through conv / upfirdn2d / bias_act / modulation -- against the reference's own
    SG2Loss / R1reg / PPLreg run on CPU (oracle/make_golden.py)."""
    from sgb200 import training
    z, meta = gold
    cfg, _, _ = _build(meta, 'cpu', channels_last=channels_last)
    tr = training.Trainer(cfg, DEV)
    _load(z, tr.G, tr.D)
    zz = torch.from_numpy(z['z']).to(DEV)
    real = torch.from_numpy(z['real']).to(DEV)
    gains = meta['gains']

    def run(name, fn):
        mod = tr.G if name.startswith('G') else tr.D
        for p in mod.parameters():
            p.grad = None
        mod.requires_grad_(True)
        fn()
        mod.requires_grad_(False)
        return mod

    _check(z, 'Gmain', 'G.', run('Gmain', lambda: tr.phase_Gmain(zz, gains['Gmain'])), 2e-4)
    _check(z, 'Dmain', 'D.', run('Dmain', lambda: tr.phase_Dmain(zz, real, gains['Dmain'])), 2e-4)
    _check(z, 'Dreg', 'D.', run('Dreg', lambda: tr.phase_Dreg(real, gains['Dreg'])), 5e-4)
    pl_noise = torch.from_numpy(z['pl_noise']).to(DEV)
    tr.pl_mean.zero_()
    _check(z, 'Greg', 'G.', run('Greg', lambda: tr.phase_Greg(zz, gains['Greg'], pl_noise=pl_noise)), 5e-4)


@pytest.mark.gpu
def test_fp16_network_against_fp32_oracle():
    """num_fp16_res=2 + conv_clamp=256 (the config-f / sg2attent numerics) at 32x32: 1e-2 against the fp32 oracle
    run with the same weights."""
    from sgb200 import training
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                               num_fp16_res=2, conv_clamp=256.0, batch_gpu=4, style_mixing_prob=0.0, noise_mode='const',
                               use_ema=False, mbstd_group_size=2)
    torch.manual_seed(0)
    G, D = training.build_networks(cfg, DEV)
    with torch.no_grad():
        for n_, p in G.named_parameters():
            if n_.endswith('noise_strength'):
                p.fill_(0.1)
    ocfg = RN.NetConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                        num_fp16_res=0, conv_clamp=256.0, d_arch='resnet', mbstd_group_size=2)
    GP = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    DP = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    zz = torch.randn(4, 64)
    with torch.no_grad():
        img = G.synthesis(G.mapping(zz.to(DEV), None, skip_w_avg_update=True), noise_mode='const')
        logits = D(img, None)
        img_o = RN.g_synthesis(GP, RN.g_mapping(GP, zz, ocfg), ocfg, noise='const')
        logits_o = RN.d_forward(DP, img.cpu().float(), ocfg)
    assert img.dtype == torch.float32
    assert_close(img, img_o, 1e-2, 'img fp16 blocks')
    assert_close(logits, logits_o, 2e-2, 'logits fp16 blocks')


@pytest.mark.gpu
def test_trainer_iteration_runs_all_phases():
    from sgb200 import training, _lib
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=512, channel_max=32, map_layers=2,
                               batch_gpu=4, mbstd_group_size=2, g_reg_interval=4, d_reg_interval=2)
    tr = training.Trainer(cfg, DEV)
    real = torch.randint(0, 256, [4, 3, 32, 32], dtype=torch.uint8, device=DEV)
    before = [p.detach().clone() for p in tr.G.parameters()]
    n0 = _lib.launch_count()
    seen = set()
    for i in range(4):
        out = tr.iteration(real)
        seen |= set(out)
        for v in out.values():
            assert torch.isfinite(v).all()
    assert seen == {'Gmain', 'Greg', 'Dmain', 'Dreg'}
    assert _lib.launch_count() > n0
    assert any((a != b.detach()).any() for a, b in zip(before, tr.G.parameters()))


@pytest.mark.gpu
def test_cuda_graph_training_iterations():
    """Each phase captured in a CUDA graph and replayed: losses stay finite, parameters move, the lazy
    regularisation schedule is honoured, and the replayed launch count is what the capture recorded."""
    import torch
    from sgb200 import training
    dev = torch.device('cuda', 0)
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                               batch_gpu=4, mbstd_group_size=2, g_reg_interval=4, d_reg_interval=2, cuda_graphs=True)
    tr = training.Trainer(cfg, dev)
    real = torch.randint(0, 256, [4, 3, 32, 32], dtype=torch.uint8, device=dev)
    p0 = [p.detach().clone() for p in tr.G.parameters()]
    seen = []
    for i in range(4):
        out = tr.iteration(real)
        torch.cuda.synchronize()
        seen.append(sorted(out))
        assert all(torch.isfinite(v).all() for v in out.values()), (i, out)
    assert seen[0] == ['Dmain', 'Dreg', 'Gmain', 'Greg'] and seen[1] == ['Dmain', 'Gmain'] and seen[2] == ['Dmain', 'Dreg', 'Gmain']
    assert tr.replayed_launches > 0
    assert any((a - b.detach()).abs().max() > 0 for a, b in zip(p0, tr.G.parameters()))
    assert all(torch.isfinite(p).all() for p in list(tr.G.parameters()) + list(tr.D.parameters()))


@pytest.mark.gpu
def test_multi_tensor_nan_to_num():
    """sgb_nan_to_num_multi == torch.nan_to_num per tensor (trainers.py:745-748), > 96 tensors, empty and odd sizes."""
    from sgb200.training import nan_to_num_
    torch.manual_seed(3)
    sizes = [0, 1, 3, 7, 512, 1000, 4097, 512 * 9 * 31] + [5 + i for i in range(120)]
    ts = []
    for i, n in enumerate(sizes):
        t = torch.randn(n, device='cuda') * 1e3
        if n > 0:
            t[torch.randint(0, n, [max(1, n // 7)], device='cuda')] = float('nan')
            t[torch.randint(0, n, [max(1, n // 9)], device='cuda')] = float('inf')
            t[torch.randint(0, n, [max(1, n // 11)], device='cuda')] = float('-inf')
        ts.append(t)
    ts.append(torch.randn(6, 5, device='cuda').t())                      # non-contiguous: falls back per tensor
    ts.append(torch.randn(9, device='cuda', dtype=torch.float16))
    ts[-1][2] = float('nan')
    want = [torch.nan_to_num(t, nan=0, posinf=1e5, neginf=-1e5) for t in ts]
    nan_to_num_(ts, nan=0, posinf=1e5, neginf=-1e5)
    for a, b in zip(ts, want):
        assert torch.equal(a, b)
